"""CUDA estimators / whitening / value loss / KL reward shaping (csrc/estimator_kernels.cuh)
through the Python mirror and the C ABI, against the reference-generated golden vectors and the CPU oracle.

Bars: the sequence recurrences (REINFORCE++ returns, GAE returns) run in the reference's own fp32 operation order and
must be BIT-EXACT; everything that passes through a reduction (group sums, masked mean / variance) is held to the
advantage tolerance of BASELINE.json's north_star, 1e-6 relative to max(1, |reference|)."""
import numpy as np
import pytest
import torch

from oracle import grpo_oracle as O

pytestmark = pytest.mark.gpu
TAGS = ("s", "m", "l")
TOL = 1e-6


@pytest.fixture(scope="module")
def st():
    import spatialthinker_b200 as st

    st.load_library()
    return st


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def t(a):
    return torch.from_numpy(np.asarray(a))


def close(got, want, tol=TOL, batch_scale=False):
    """|got - want| <= tol * max(1, |want|) element by element. ``batch_scale``: relative to the largest reference value
    instead - for per-token ("dense") rewards, where a sequence score is a rounded fp32 sum of many terms and a group
    baseline carries the rounding of its largest member into members whose own advantage is small."""
    got, want = got.detach().float().cpu(), torch.as_tensor(want).float()
    err = (got - want).abs()
    scale = want.abs().max().clamp_min(1.0) if batch_scale else want.abs().clamp_min(1.0)
    assert bool((err <= tol * scale).all()), float(err.max())


def uid_of(g, tag):
    return np.array([str(u) for u in g[f"{tag}_uid"]], dtype=object)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("rname", ["sparse", "dense"])
def test_outcome_estimators_golden(st, dev, golden, tag, rname):
    g = golden("estimators")
    rew, mask = t(g[f"{tag}_{rname}"]).to(dev), t(g[f"{tag}_mask"]).to(dev)
    adv, ret = st.compute_rloo_outcome_advantage(rew, mask, uid_of(g, tag))
    assert adv is ret and adv.dtype == torch.float32
    close(adv, g[f"{tag}_{rname}_rloo"], batch_scale=rname == "dense")
    adv, ret = st.compute_remax_outcome_advantage(rew, t(g[f"{tag}_baselines"]).to(dev), mask)
    assert adv is ret
    close(adv, g[f"{tag}_{rname}_remax"], batch_scale=rname == "dense")


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("rname", ["sparse", "dense"])
def test_recurrent_estimators_golden(st, dev, golden, tag, rname):
    g = golden("estimators")
    rew, mask, values = t(g[f"{tag}_{rname}"]).to(dev), t(g[f"{tag}_mask"]).to(dev), t(g[f"{tag}_values"]).to(dev)
    for gamma in (1.0, 0.97):
        adv, ret = st.compute_reinforce_plus_plus_outcome_advantage(rew, mask, gamma)
        np.testing.assert_array_equal(ret.cpu().numpy(), g[f"{tag}_{rname}_rpp_ret_{gamma}"])  # bit-exact recurrence
        close(adv, g[f"{tag}_{rname}_rpp_adv_{gamma}"])
    for gamma, lam in ((1.0, 1.0), (0.99, 0.95)):
        adv, ret = st.compute_gae_advantage_return(rew, values, mask, gamma, lam)
        np.testing.assert_array_equal(ret.cpu().numpy(), g[f"{tag}_{rname}_gae_ret_{gamma}_{lam}"])
        close(adv, g[f"{tag}_{rname}_gae_adv_{gamma}_{lam}"])


@pytest.mark.parametrize("tag", TAGS)
def test_whiten_and_value_loss_golden(st, dev, golden, tag):
    g = golden("estimators")
    mask, values = t(g[f"{tag}_mask"]).to(dev), t(g[f"{tag}_values"]).to(dev)
    close(st.masked_var(values, mask), g[f"{tag}_var"][0])
    close(st.masked_var(values, mask, unbiased=False), g[f"{tag}_var"][1])
    close(st.masked_whiten(values, mask), g[f"{tag}_whiten"])
    for m in (mask.float(), mask.bool()):
        close(st.masked_whiten(values, m), g[f"{tag}_whiten"])
    vp = t(g[f"{tag}_vpreds"]).to(dev).requires_grad_(True)
    loss, frac = st.compute_value_loss(vp, t(g[f"{tag}_returns"]).to(dev), values, mask, 0.5)
    assert loss.dim() == 0 and frac.dim() == 0
    (3.0 * loss).backward()
    close(loss, g[f"{tag}_vf"][0])
    close(frac, g[f"{tag}_vf"][1])
    close(vp.grad / 3.0, g[f"{tag}_vf_grad"])


@pytest.mark.parametrize("tag", TAGS)
def test_kl_reward_shaping_golden(st, dev, golden, tag):
    g = golden("estimators")
    mask = t(g[f"{tag}_mask"]).to(dev)
    old, ref, sparse = t(g[f"{tag}_old"]).to(dev), t(g[f"{tag}_ref"]).to(dev), t(g[f"{tag}_sparse"]).to(dev)
    for mode in O.KL_MODES:
        data = st.TensorBatch({"token_level_scores": sparse, "old_log_probs": old, "ref_log_probs": ref, "response_mask": mask})
        ctl = st.core_algos.AdaptiveKLController(init_kl_coef=0.05, target_kl=0.1, horizon=1000.0)
        out, metrics = st.apply_kl_penalty(data, ctl, kl_penalty=mode)
        assert out is data and metrics["critic/kl_coef"] == 0.05
        close(data.batch["token_level_rewards"], g[f"{tag}_klrew_{mode}"])
        want = float(g[f"{tag}_klcur_{mode}"][0])
        assert abs(metrics["critic/kl"] - want) <= 1e-6 * max(1.0, abs(want))
        assert ctl.kl_coef != 0.05  # the controller saw the measurement
    data = st.TensorBatch({"token_level_scores": sparse, "response_mask": mask})
    _, metrics = st.apply_kl_penalty(data, st.core_algos.FixedKLController(0.05))
    assert torch.equal(data.batch["token_level_rewards"], sparse) and metrics["critic/kl"] == 0.0
    np.testing.assert_array_equal(st.compute_rewards(sparse, old, ref, 0.05).cpu().numpy(), g[f"{tag}_compute_rewards"])
    with pytest.raises(NotImplementedError):
        st.ray_trainer.kl_penalty_rewards(sparse, old, ref, mask, 0.05, "full")


def test_compute_advantage_dispatch(st, dev, golden):
    g = golden("estimators")
    tag = "m"
    batch = {k: t(g[f"{tag}_{s}"]).to(dev) for k, s in (("token_level_rewards", "dense"), ("response_mask", "mask"),
                                                       ("values", "values"), ("reward_baselines", "baselines"))}
    E = st.AdvantageEstimator
    want = {E.RLOO: "rloo", E.REMAX: "remax", E.REINFORCE_PLUS_PLUS: "rpp_adv_0.97", E.GAE: "gae_adv_0.99_0.95"}
    for est, key in want.items():
        data = st.TensorBatch(dict(batch), {"uid": uid_of(g, tag)})
        st.compute_advantage(data, est, gamma=0.97 if est == E.REINFORCE_PLUS_PLUS else 0.99, lam=0.95)
        close(data.batch["advantages"], g[f"{tag}_dense_{key}"], batch_scale=True)
        assert data.batch["returns"].shape == data.batch["advantages"].shape
    data = st.TensorBatch(dict(batch), {"uid": uid_of(g, tag)})
    st.compute_advantage(data, "grpo")  # the enum is a str: plain strings dispatch too
    want_adv, _ = O.compute_grpo_outcome_advantage(batch["token_level_rewards"].cpu().clone(), batch["response_mask"].cpu(), uid_of(g, tag))
    close(data.batch["advantages"], want_adv, 2e-6)
    assert data.batch["returns"] is data.batch["advantages"]
    with pytest.raises(NotImplementedError):
        st.compute_advantage(data, "vtrace")


def test_full_size_vs_oracle_and_edges(st, dev):
    # config C5's shape class: ragged masks, n = 16, long responses (T = 4096); 1024 sequences keep the oracle's Python
    # loops to a few seconds
    roll = O.synth_rollout(1024, 4096, 151936, 16, seed=21, ragged=True)
    mask = roll["response_mask"]
    g = torch.Generator().manual_seed(3)
    dense = torch.randn(mask.shape, generator=g) * mask * 0.05 + roll["token_level_rewards"]
    values = torch.randn(mask.shape, generator=g)
    d_mask, d_dense, d_values = mask.to(dev), dense.to(dev), values.to(dev)
    # per-token rewards make the sequence score a 4096-term sum: two fp32 summation orders differ by more than 1e-6, so
    # the yardstick here is the oracle run in fp64 (the kernel accumulates in fp64 and rounds once)
    want, _ = O.compute_rloo_outcome_advantage(dense.double(), mask, roll["uid"])
    close(st.compute_rloo_outcome_advantage(d_dense, d_mask, roll["uid"])[0], want)
    want, _ = O.compute_grpo_outcome_advantage(dense.double(), mask, roll["uid"])
    close(st.compute_grpo_outcome_advantage(d_dense, d_mask, roll["uid"])[0], want, 2e-6)
    wadv, wret = O.compute_reinforce_plus_plus_outcome_advantage(dense.clone(), mask, 0.999)
    adv, ret = st.compute_reinforce_plus_plus_outcome_advantage(d_dense, d_mask, 0.999)
    assert torch.equal(ret.cpu(), wret)
    close(adv, wadv, 2e-6)
    wadv, wret = O.compute_gae_advantage_return(dense.clone(), values, mask, 0.999, 0.95)
    adv, ret = st.compute_gae_advantage_return(d_dense, d_values, d_mask, 0.999, 0.95)
    assert torch.equal(ret.cpu(), wret)
    close(adv, wadv, 2e-6)
    # whitening is idempotent up to rounding, has zero masked mean and unit masked variance
    w = st.masked_whiten(d_values, d_mask)
    assert abs(float(st.masked_mean(w, d_mask))) < 1e-5 and abs(float(st.masked_var(w, d_mask)) - 1.0) < 1e-4
    close(st.masked_whiten(w, d_mask), w.cpu(), 1e-4)
    # edges: one valid element keeps the biased variance (0), an empty mask gives 0; a group of one asserts like the
    # reference; T not a multiple of the 32-token scan tile; a single sequence
    x = torch.tensor([[1.0, 2.0, 4.0]], device=dev)
    assert float(st.masked_var(x, torch.tensor([[0, 1, 0]], device=dev))) == 0.0
    assert float(st.masked_var(x, torch.zeros(1, 3, device=dev))) == 0.0
    with pytest.raises(AssertionError, match="RLOO needs rollout.n > 1."):
        st.compute_rloo_outcome_advantage(torch.ones(3, 2, device=dev), torch.ones(3, 2, device=dev),
                                          np.array(["a", "a", "b"], dtype=object))
    for bsz, tl in ((1, 1), (1, 33), (33, 31), (130, 65)):
        m = (torch.rand(bsz, tl, generator=g) < 0.8).long()
        m[:, 0] = 1
        r, v = torch.randn(bsz, tl, generator=g), torch.randn(bsz, tl, generator=g)
        _, wret = O.compute_reinforce_plus_plus_outcome_advantage(r.clone(), m, 0.9)
        assert torch.equal(st.compute_reinforce_plus_plus_outcome_advantage(r.to(dev), m.to(dev), 0.9)[1].cpu(), wret)
        _, wret = O.compute_gae_advantage_return(r.clone(), v, m, 0.9, 0.8)
        assert torch.equal(st.compute_gae_advantage_return(r.to(dev), v.to(dev), m.to(dev), 0.9, 0.8)[1].cpu(), wret)
    with pytest.raises(RuntimeError):
        st.masked_whiten(torch.ones(2, 2), torch.ones(2, 2))  # CPU tensors: no fallback


def test_experience_pass_matches_oracle(st, dev):
    """ray_trainer.py:633-663 on the device: old / ref log-probs through the fused head, KL-shaped rewards, RLOO
    advantages - against the oracle's composition of the same steps."""
    bsz, tl, h, v, n = 16, 40, 128, 4096, 4
    hid, w = O.synth_head(bsz * tl, h, v, seed=5, sigma_w=0.1)
    hid_ref, w_ref = O.synth_head(bsz * tl, h, v, seed=6, sigma_w=0.1)
    hid_ref = (0.9 * hid.float() + 0.1 * hid_ref.float()).to(torch.bfloat16)
    w_ref = (0.9 * w.float() + 0.1 * w_ref.float()).to(torch.bfloat16)
    roll = O.synth_rollout(bsz, tl, v, n, seed=5, ragged=True)
    mask, labels = roll["response_mask"], roll["responses"]
    want_old, _ = O.lm_head_log_probs(hid.view(bsz, tl, h), w, labels, 0.9)
    want_ref, _ = O.lm_head_log_probs(hid_ref.view(bsz, tl, h), w_ref, labels, 0.9)
    want_rew, want_kl = O.kl_penalty_rewards(roll["token_level_rewards"], want_old, want_ref, mask, 0.02, "low_var_kl")
    want_adv, _ = O.compute_rloo_outcome_advantage(want_rew.clone(), mask, roll["uid"])

    cfg = st.ActorConfig(micro_batch_size_per_device_for_experience=4)
    actor = st.DataParallelPPOActor(cfg, w.to(dev))
    ref_actor = st.DataParallelPPOActor(cfg, w_ref.to(dev), hidden_fn=lambda mb: mb["ref_hidden_states"])
    data = st.TensorBatch({"hidden_states": hid.view(bsz, tl, h).to(dev), "ref_hidden_states": hid_ref.view(bsz, tl, h).to(dev),
                           "responses": labels.to(dev), "response_mask": mask.to(dev),
                           "token_level_scores": roll["token_level_rewards"].to(dev)},
                          {"uid": roll["uid"]}, {"temperature": 0.9})
    ctl = st.core_algos.FixedKLController(0.02)
    out, metrics = st.experience_pass(data, actor, ref_actor, adv_estimator="rloo", use_kl_loss=False, kl_ctrl=ctl,
                                      kl_penalty="low_var_kl")
    valid = mask.bool()
    assert float((out.batch["old_log_probs"].cpu() - want_old)[valid].abs().max()) < 2e-3
    assert float((out.batch["ref_log_probs"].cpu() - want_ref)[valid].abs().max()) < 2e-3
    # rewards and advantages inherit the 2e-3 log-prob tolerance through the KL term (coefficient 0.02)
    assert float((out.batch["token_level_rewards"].cpu() - want_rew).abs().max()) < 1e-3
    assert float((out.batch["advantages"].cpu() - want_adv).abs().max()) < 2e-2
    assert abs(metrics["critic/kl"] - want_kl) < 1e-3 and metrics["critic/kl_coef"] == 0.02
    # KL as a loss term (the shipped GRPO configuration): rewards are the scores, untouched
    out, metrics = st.experience_pass(data, actor, ref_actor, adv_estimator="grpo", use_kl_loss=True)
    assert out.batch["token_level_rewards"] is out.batch["token_level_scores"] and metrics == {}
    want_adv, _ = O.compute_grpo_outcome_advantage(roll["token_level_rewards"].clone(), mask, roll["uid"])
    close(out.batch["advantages"], want_adv)


def test_packed_response_rows_match_pad_and_slice(st, dev):
    """padding_free layout (dp_actor.py:85-139): the rows handed to the fused head must be the rows whose log-probs the
    reference keeps after pad_input + [:, -T-1:-1], and the gradient must land on exactly those packed rows."""
    from spatialthinker_b200 import hf_hook

    g = torch.Generator().manual_seed(4)
    bsz, prompt, t_len, h = 5, 11, 9, 64
    seqlen = prompt + t_len
    mask = torch.zeros(bsz, seqlen, dtype=torch.int64)
    for i in range(bsz):  # left-padded prompts, right-padded responses, as the reference's collate produces
        p_len = int(torch.randint(1, prompt + 1, (1,), generator=g))
        r_len = int(torch.randint(1, t_len + 1, (1,), generator=g))
        mask[i, prompt - p_len: prompt + r_len] = 1
    nnz = int(mask.sum())
    packed = torch.randn(nnz, h, generator=g).to(torch.bfloat16)
    # reference semantics in plain torch: pad back to (bsz, seqlen), slice
    padded = torch.zeros(bsz * seqlen, h, dtype=torch.bfloat16)
    padded[mask.view(-1).bool()] = packed
    want = padded.view(bsz, seqlen, h)[:, -t_len - 1: -1]
    x = packed.to(dev).requires_grad_(True)
    got = hf_hook.packed_response_hidden_states(x.unsqueeze(0), mask.to(dev), t_len)
    assert got.shape == (bsz, t_len, h) and torch.equal(got.cpu(), want)
    # every valid response slot found its row
    resp_mask = mask[:, -t_len:].bool()
    assert bool((want.float().abs().sum(-1) > 0)[resp_mask].all())
    up = torch.randn(bsz, t_len, h, generator=g).to(torch.bfloat16)
    got.backward(up.to(dev))
    want_grad = torch.zeros(bsz, seqlen, h, dtype=torch.bfloat16)
    want_grad[:, -t_len - 1: -1] = up
    want_grad = want_grad.view(-1, h)[mask.view(-1).bool()]
    assert torch.equal(x.grad.cpu(), want_grad)
