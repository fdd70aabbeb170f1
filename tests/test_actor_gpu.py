"""GPU tests of the actor loop's round-2 additions: the optimizer step over every parameter (global-norm clip, tied
weights, non-finite skip - verl/workers/actor/dp_actor.py:155-167), token-balanced micro-batching
(verl/utils/seqlen_balancing.py:222-255), the run-to-run reproducible mode and the gradient passes of
csrc/grad_kernels.cuh. Checked against the CPU oracle / plain torch autograd on fp32 copies."""
import math
import warnings

import numpy as np
import pytest
import torch

from oracle import grpo_oracle as O

pytestmark = pytest.mark.gpu
TOL_LOGP, TOL_REL = 2e-3, 1e-2
CLIP = dict(clip_ratio_low=0.2, clip_ratio_high=0.3, clip_ratio_dual=3.0)


@pytest.fixture(scope="module")
def st():
    import spatialthinker_b200 as st

    st.load_library()
    return st


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _inputs(bsz, tl, h, v, n, sigma, seed, ragged=True):
    hid, w = O.synth_head(bsz * tl, h, v, seed=seed, sigma_w=sigma)
    hid = hid.view(bsz, tl, h)
    roll = O.synth_rollout(bsz, tl, v, n, seed=seed, ragged=ragged)
    logp, _ = O.lm_head_log_probs(hid, w, roll["responses"])
    adv, _ = O.compute_grpo_outcome_advantage(roll["token_level_rewards"].clone(), roll["response_mask"], roll["uid"])
    return {"hidden": hid, "weight": w, "labels": roll["responses"], "mask": roll["response_mask"], "adv": adv,
            "old": O.perturbed_log_probs(logp, seed=seed + 1, outlier_frac=0.03),
            "ref": O.perturbed_log_probs(logp, seed=seed + 2, outlier_frac=0.03)}


# ================================================================================================ gradient passes
@pytest.mark.parametrize("n", [1, 7, 4096, 1000003])
def test_grad_sumsq_and_scale_cast(st, dev, n):
    from spatialthinker_b200.dp_actor import grad_scale_cast, grad_sumsq

    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, generator=g)
    xd = x.to(dev)
    want = float(x.double().square().sum())
    got = grad_sumsq(xd)
    assert got.dtype == torch.float64 and abs(float(got) - want) <= 1e-6 * want
    assert torch.equal(xd.cpu(), x)  # untouched
    acc = grad_sumsq(xd, out=got.clone())  # accumulate: twice the sum
    assert abs(float(acc) - 2 * want) <= 2e-6 * want
    assert float(grad_sumsq(xd)) == float(got)  # fixed reduction order: bit-identical
    out = torch.empty(n, dtype=torch.bfloat16, device=dev)
    scale = torch.tensor([0.37], device=dev)
    grad_scale_cast(xd, scale, out)
    assert torch.equal(out.cpu(), (x * 0.37).to(torch.bfloat16))
    grad_scale_cast(xd, 2.0, out, zero_after=True)
    assert torch.equal(out.cpu(), (x * 2.0).to(torch.bfloat16)) and float(xd.abs().max()) == 0.0
    y = x.to(dev)
    s2 = grad_sumsq(y, zero_after=True)
    assert float(s2) == float(got) and float(y.abs().max()) == 0.0


# ================================================================================================ optimizer step
def test_optimizer_step_global_clip_with_tied_embedding(st, dev):
    """A body whose embedding is TIED to the head weight (the 3B checkpoints) plus a linear layer, all held by the
    optimizer: after update_policy the weights must equal what torch autograd + clip_grad_norm_ over ALL parameters + SGD
    give on fp32 copies (dp_actor.py:155-167), and actor/grad_norm is the global norm."""
    bsz, tl, h, v = 8, 24, 64, 1024
    g = torch.Generator().manual_seed(5)
    w0 = (0.3 * torch.randn(v, h, generator=g)).to(torch.bfloat16)
    a0 = (torch.eye(h) + 0.05 * torch.randn(h, h, generator=g)).to(torch.bfloat16)
    ids = torch.randint(0, v, (bsz, tl), generator=g)
    roll = O.synth_rollout(bsz, tl, v, 4, seed=5, ragged=True)
    labels, mask = roll["responses"], roll["response_mask"]
    adv = 4.0 * torch.randn(bsz, 1, generator=g).expand(bsz, tl).contiguous()  # large enough for the clip to bind
    with torch.no_grad():
        hid0 = (w0.float()[ids] @ a0.float().t()).to(torch.bfloat16)
        lp0, _ = O.lm_head_log_probs(hid0, w0, labels)
    old = O.perturbed_log_probs(lp0, seed=6, outlier_frac=0.03)
    max_norm, lr = 0.05, 0.5

    # ---- reference: autograd on fp32 leaves holding the bf16 values, two micro-batches of 4, GA = 2
    w_ref = w0.float().clone().requires_grad_(True)
    a_ref = a0.float().clone().requires_grad_(True)
    for sl in (slice(0, 4), slice(4, 8)):
        hid = (w_ref[ids[sl]] @ a_ref.t()).to(torch.bfloat16).float()  # the body emits bf16 hidden states (the casts are
        # differentiable: the gradient passes straight through)
        z = hid @ w_ref.t()
        lp = O.log_probs_from_logits(z, labels[sl])
        loss, _ = O.micro_batch_loss(lp, old[sl], adv[sl], mask[sl], None, grad_accum=2.0, **CLIP)
        loss.backward()
    want_norm = float(torch.nn.utils.clip_grad_norm_([w_ref, a_ref], max_norm))
    assert want_norm > max_norm  # the clip binds: w_ref.grad / a_ref.grad now hold the CLIPPED gradients

    # ---- ours
    weight = torch.nn.Parameter(w0.to(dev).clone())
    a_par = torch.nn.Parameter(a0.to(dev).clone())
    opt = torch.optim.SGD([weight, a_par], lr=lr)

    def hidden_fn(mb):
        return torch.nn.functional.linear(torch.nn.functional.embedding(mb["input_ids"], weight), a_par)

    seen = {}  # the gradients the optimizer is handed (the bf16 weights cannot resolve an update of norm 0.05 * lr)
    opt.register_step_pre_hook(lambda o, a, k: seen.update(w=weight.grad.detach().clone(), a=a_par.grad.detach().clone()))
    cfg = st.ActorConfig(global_batch_size_per_device=8, micro_batch_size_per_device_for_update=4, max_grad_norm=max_norm)
    actor = st.DataParallelPPOActor(cfg, weight, actor_optimizer=opt, hidden_fn=hidden_fn)
    data = st.TensorBatch({"input_ids": ids.to(dev), "responses": labels.to(dev), "response_mask": mask.to(dev),
                           "old_log_probs": old.to(dev), "advantages": adv.to(dev)}, meta_info={"temperature": 1.0})
    met = actor.update_policy(data)
    assert abs(met["actor/grad_norm"][0] - want_norm) <= 2e-2 * want_norm
    assert rel(seen["w"], w_ref.grad) < 2e-2  # head + tied-embedding contributions, clipped by the global norm
    assert rel(seen["a"], a_ref.grad) < 2e-2
    clipped = math.sqrt(float(seen["w"].float().square().sum() + seen["a"].float().square().sum()))
    assert abs(clipped - max_norm) <= 2e-2 * max_norm
    assert weight.grad is None and a_par.grad is None  # zero_grad: nothing piles up across steps
    # a second update starts from clean gradients
    met2 = actor.update_policy(data)
    assert math.isfinite(met2["actor/grad_norm"][0])


def test_optimizer_step_skips_on_non_finite_norm(st, dev, capsys):
    """dp_actor.py:161-164: a non-finite gradient norm prints and SKIPS optimizer.step(); AdamW's moments, step count and
    the weights are untouched, and the accumulator is clean for the next step."""
    bsz, tl, h, v = 4, 16, 64, 1024
    x = _inputs(bsz, tl, h, v, 4, 0.1, seed=17, ragged=False)
    weight = torch.nn.Parameter(x["weight"].to(dev).clone())
    opt = torch.optim.AdamW([weight], lr=1e-2, weight_decay=0.1)
    cfg = st.ActorConfig(global_batch_size_per_device=4, micro_batch_size_per_device_for_update=4)
    actor = st.DataParallelPPOActor(cfg, weight, actor_optimizer=opt)
    base = {"hidden_states": x["hidden"].to(dev), "responses": x["labels"].to(dev), "response_mask": x["mask"].to(dev),
            "old_log_probs": x["old"].to(dev)}
    good = st.TensorBatch({**base, "advantages": x["adv"].to(dev)}, meta_info={"temperature": 1.0})
    actor.update_policy(good)  # one ordinary step so that AdamW has state
    w1 = weight.detach().clone()
    state1 = {k: (val.clone() if torch.is_tensor(val) else val) for k, val in opt.state[weight].items()}
    bad_adv = x["adv"].clone()
    bad_adv[0, 0] = float("inf")
    met = actor.update_policy(st.TensorBatch({**base, "advantages": bad_adv.to(dev)}, meta_info={"temperature": 1.0}))
    assert not math.isfinite(met["actor/grad_norm"][0])
    assert "Gradient norm is not finite. Skip update." in capsys.readouterr().out
    assert torch.equal(weight.detach(), w1)
    for k, val in opt.state[weight].items():
        assert (torch.equal(val, state1[k]) if torch.is_tensor(val) else val == state1[k]), k
    assert float(actor.dweight.abs().max()) == 0.0
    met3 = actor.update_policy(good)
    assert math.isfinite(met3["actor/grad_norm"][0]) and not torch.equal(weight.detach(), w1)


# ================================================================================================ dynamic micro-batches
def test_update_policy_dynamic_bsz_matches_oracle(st, dev):
    """use_dynamic_bsz: micro-batches by token count (Karmarkar-Karp, seqlen_balancing.py:222-255), each weighted by its
    share of the mini-batch's sequences. Checked against the oracle run over the SAME partitions."""
    from spatialthinker_b200.sharding import rearrange_micro_batches

    bsz, tl, h, v = 16, 64, 128, 4096
    x = _inputs(bsz, tl, h, v, 4, 0.1, seed=23, ragged=True)
    lens = x["mask"].sum(-1).tolist()
    max_tokens = 160
    cfg = st.ActorConfig(global_batch_size_per_device=8, micro_batch_size_per_device_for_update=2, use_dynamic_bsz=True,
                         max_token_len_per_micro_batch=max_tokens, use_kl_loss=True, kl_penalty="low_var_kl", kl_coef=0.01)
    actor = st.DataParallelPPOActor(cfg, x["weight"].to(dev))
    dws = []
    orig = actor._optimizer_step
    actor._optimizer_step = lambda: (dws.append(actor.dweight.clone()), orig())[1]
    data = st.TensorBatch({"hidden_states": x["hidden"].to(dev), "responses": x["labels"].to(dev),
                           "response_mask": x["mask"].to(dev), "old_log_probs": x["old"].to(dev),
                           "advantages": x["adv"].to(dev), "ref_log_probs": x["ref"].to(dev)}, meta_info={"temperature": 1.0})
    met = actor.update_policy(data)
    want_pg, k = [], 0
    for step, s0 in enumerate((0, 8)):
        parts = rearrange_micro_batches(lens[s0:s0 + 8], max_tokens)
        assert len(parts) == -(-sum(lens[s0:s0 + 8]) // max_tokens) and sorted(i for p in parts for i in p) == list(range(8))
        sums = [sum(lens[s0 + i] for i in p) for p in parts]
        assert max(sums) - min(sums) <= max(lens[s0:s0 + 8])  # balanced
        dw = torch.zeros(v, h)
        for p in parts:
            idx = torch.tensor([s0 + i for i in p])
            res = O.fused_loss_reference(x["hidden"][idx], x["weight"], x["labels"][idx], x["old"][idx], x["adv"][idx],
                                         x["mask"][idx], x["ref"][idx], kl_penalty="low_var_kl", kl_coef=0.01,
                                         grad_accum=8.0 / len(p), **CLIP)
            dw += res["dweight"]
            want_pg.append(float(res["metrics"]["actor/pg_loss"]))
            k += 1
        assert rel(dws[step], dw) < TOL_REL
    assert len(met["actor/pg_loss"]) == k and len(met["actor/grad_norm"]) == 2
    np.testing.assert_allclose(met["actor/pg_loss"], want_pg, rtol=TOL_REL, atol=1e-5)


# ================================================================================================ reproducible mode
def test_deterministic_mode_is_bit_reproducible(st, dev):
    """Option "deterministic" (ActorConfig.deterministic): no split-K, one-hot rows of dW summed in row order. Labels
    are drawn from 64 tokens only, so every dW row receives hundreds of one-hot contributions; rows x H are chosen so that
    the default schedule takes both the dHidden split path and the dW split-K tail."""
    from spatialthinker_b200 import _lib

    lib = _lib.load()
    rows, h, v = 4096 - 33, 512, 8192 + 72
    g = torch.Generator().manual_seed(9)
    hid = torch.randn(rows, h, generator=g).to(torch.bfloat16)
    w = (0.05 * torch.randn(v, h, generator=g)).to(torch.bfloat16)
    lab = torch.randint(0, 64, (rows,), generator=g)
    lab[::7] = torch.randint(0, v, (len(lab[::7]),), generator=g)
    mask = (torch.rand(rows, generator=g) > 0.2).long()
    adv = torch.randn(rows, generator=g)
    lp, _ = O.lm_head_log_probs(hid, w, lab)
    old = O.perturbed_log_probs(lp, seed=3, outlier_frac=0.03)
    want = O.fused_loss_reference(hid, w, lab, old, adv, mask, None, grad_accum=2.0, **CLIP)
    args = [t_.to(dev) for t_ in (hid, w, lab, old, adv)]

    def run():
        res = st.grpo_micro_batch_step(*args, None, mask.to(dev), kl_penalty=None, grad_accum=2.0, **CLIP)
        torch.cuda.synchronize()
        return res

    _lib.check(lib.grpo_set_option(b"deterministic", 1), "set_option")
    try:
        runs = [run() for _ in range(3)]
    finally:
        _lib.check(lib.grpo_set_option(b"deterministic", 0), "set_option")
    for r in runs[1:]:
        assert torch.equal(r["dweight"], runs[0]["dweight"])
        assert torch.equal(r["dhidden"], runs[0]["dhidden"])
        assert torch.equal(r["log_probs"], runs[0]["log_probs"])
    assert rel(runs[0]["dweight"], want["dweight"]) < TOL_REL and rel(runs[0]["dhidden"], want["dhidden"]) < TOL_REL
    # the default (atomics + split-K) agrees with it to fp32 summation order
    fast = run()
    assert rel(fast["dweight"], runs[0]["dweight"]) < 1e-4 and rel(fast["dhidden"], runs[0]["dhidden"]) < 2e-3
    # and through the actor switch
    cfg = st.ActorConfig(global_batch_size_per_device=1, micro_batch_size_per_device_for_update=1, deterministic=True)
    try:
        actor = st.DataParallelPPOActor(cfg, args[1])
        data = st.TensorBatch({"hidden_states": args[0][None], "responses": args[2][None], "response_mask": mask.to(dev)[None],
                               "old_log_probs": args[3][None], "advantages": args[4][None]}, meta_info={"temperature": 1.0})
        grabbed = []
        orig = actor._optimizer_step
        actor._optimizer_step = lambda: (grabbed.append(actor.dweight.clone()), orig())[1]
        m1 = actor.update_policy(data)
        m2 = actor.update_policy(data)
        assert torch.equal(grabbed[0], grabbed[1]) and m1["actor/grad_norm"] == m2["actor/grad_norm"]
    finally:
        _lib.check(lib.grpo_set_option(b"deterministic", 0), "set_option")


# ================================================================================================ saturation count
def test_saturated_tokens_are_counted(st, dev):
    """GRPO_MET_SATURATED: unmasked tokens whose log-probability is below -69.3 (where the label-referenced softmax's
    exp2 clamp may bind) are counted and surfaced by the actor as a warning + ``actor/saturated_tokens``."""
    from spatialthinker_b200 import _lib

    bsz, tl, h, v = 2, 32, 128, 4096
    g = torch.Generator().manual_seed(7)
    hid = torch.randn(bsz, tl, h, generator=g).to(torch.bfloat16)
    w = (0.45 * torch.randn(v, h, generator=g)).to(torch.bfloat16)
    w[7] = (-8.0 * hid[0, 0].float() / hid[0, 0].float().norm()).to(torch.bfloat16)
    z = hid.float() @ w.float().t()
    labels = z.argmax(-1)
    labels[0, 0] = 7
    assert float(z[0, 0].max() - z[0, 0, 7]) > 80
    mask = torch.ones(bsz, tl, dtype=torch.int64)
    old = torch.full((bsz, tl), -1.0)
    adv = torch.ones(bsz, tl)
    res = st.grpo_micro_batch_step(hid.to(dev), w.to(dev), labels.to(dev), old.to(dev), adv.to(dev), None, mask.to(dev),
                                   kl_penalty=None)
    assert float(res["metrics"][_lib.MET_SATURATED]) == 1.0
    mask[0, 0] = 0  # masked: not counted
    res = st.grpo_micro_batch_step(hid.to(dev), w.to(dev), labels.to(dev), old.to(dev), adv.to(dev), None, mask.to(dev),
                                   kl_penalty=None)
    assert float(res["metrics"][_lib.MET_SATURATED]) == 0.0
    mask[0, 0] = 1
    cfg = st.ActorConfig(global_batch_size_per_device=2, micro_batch_size_per_device_for_update=2)
    actor = st.DataParallelPPOActor(cfg, w.to(dev))
    data = st.TensorBatch({"hidden_states": hid.to(dev), "responses": labels.to(dev), "response_mask": mask.to(dev),
                           "old_log_probs": old.to(dev), "advantages": adv.to(dev)}, meta_info={"temperature": 1.0})
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        met = actor.update_policy(data)
    assert met["actor/saturated_tokens"] == 1.0 and any("-69.3" in str(c.message) for c in caught)


# ================================================================================================ deferred dW: lazy workspace
def test_deferred_dw_workspace_is_lazy_and_releasable(st, dev):
    from spatialthinker_b200.fused import DeferredDW

    w = torch.zeros(1024, 64, dtype=torch.bfloat16, device=dev)
    dw = torch.zeros(1024, 64, dtype=torch.float32, device=dev)
    d = DeferredDW(w, dw)
    assert d.workspace is None and d.capacity % 512 == 0
    assert d.reserve(d.capacity + 1) is None and d.workspace is None  # too large: ordinary path, still nothing allocated
    assert d.reserve(100) == 0 and d.workspace is not None
    d.next_row0 = d.total_rows = d.pending = 0  # nothing was launched into the slot
    d.release()
    assert d.workspace is None


# ================================================================================================ host-resident batches
@pytest.mark.parametrize("dynamic", [False, True])
def test_update_policy_streams_from_pinned_host(st, dev, dynamic):
    """A batch in pinned host memory is streamed micro-batch by micro-batch (double-buffered copy stream): same kernels
    on the same rows as the device-resident batch, so - in the reproducible mode - dW and dHidden are identical bit for
    bit and the metrics agree to the rounding of their fp64 atomics."""
    from spatialthinker_b200 import _lib

    bsz, tl, h, v = 16, 48, 128, 4096
    x = _inputs(bsz, tl, h, v, 4, 0.1, seed=33, ragged=True)
    tensors = {"hidden_states": x["hidden"], "responses": x["labels"], "response_mask": x["mask"],
               "old_log_probs": x["old"], "advantages": x["adv"], "ref_log_probs": x["ref"]}
    cfg = st.ActorConfig(global_batch_size_per_device=8, micro_batch_size_per_device_for_update=2, use_dynamic_bsz=dynamic,
                         max_token_len_per_micro_batch=120, use_kl_loss=True, kl_penalty="low_var_kl", kl_coef=0.01,
                         deterministic=True)
    results = []
    try:
        for where in ("device", "host"):
            batch = {k: (t_.to(dev) if where == "device" else t_.clone().pin_memory()) for k, t_ in tensors.items()}
            actor = st.DataParallelPPOActor(cfg, x["weight"].to(dev))
            dws = []
            orig = actor._optimizer_step
            actor._optimizer_step = lambda a=actor, o=orig, d=dws: (d.append(a.dweight.clone()), o())[1]
            met = actor.update_policy(st.TensorBatch(batch, meta_info={"temperature": 1.0}))
            torch.cuda.synchronize()
            results.append((met, dws, torch.cat([d.reshape(-1, h) for d in actor.last_dhidden])))
    finally:
        _lib.check(_lib.load().grpo_set_option(b"deterministic", 0), "set_option")
    (m0, d0, h0), (m1, d1, h1) = results
    assert m0.keys() == m1.keys() and len(d0) == len(d1) == 2
    for key in m0:
        np.testing.assert_allclose(m0[key], m1[key], rtol=1e-6, atol=1e-9)
    for a, b in zip(d0, d1):
        assert torch.equal(a, b)
    assert torch.equal(h0, h1)


# ================================================================================================ speed-aware shards
@pytest.mark.parametrize("dynamic", [False, True])
def test_uneven_rank_shards_keep_the_gradient(st, dev, dynamic):
    """ActorConfig.loss_scale_batch_size: two ranks holding 10 and 6 sequences of a 16-sequence mini-batch (speed-aware
    shards) scale their losses for the nominal 8 - the mean of their dW over ranks is the gradient of the equal 8 + 8 split
    and of the oracle's global loss. Dense masks: every token then weighs the same whichever rank and micro-batch holds it."""
    bsz, tl, h, v = 16, 32, 128, 2048
    x = _inputs(bsz, tl, h, v, 4, 0.1, seed=61, ragged=False)
    w = x["weight"].to(dev)

    def rank_dw(rows, mini, nominal):
        cfg = st.ActorConfig(global_batch_size_per_device=mini, loss_scale_batch_size=nominal,
                             micro_batch_size_per_device_for_update=2, use_dynamic_bsz=dynamic,
                             max_token_len_per_micro_batch=3 * tl, use_kl_loss=True, kl_penalty="low_var_kl", kl_coef=0.01)
        actor = st.DataParallelPPOActor(cfg, w)
        grabbed = []
        orig = actor._optimizer_step
        actor._optimizer_step = lambda: (grabbed.append(actor.dweight.clone()), orig())[1]
        sl = slice(*rows)
        data = st.TensorBatch({"hidden_states": x["hidden"][sl].to(dev), "responses": x["labels"][sl].to(dev),
                               "response_mask": x["mask"][sl].to(dev), "old_log_probs": x["old"][sl].to(dev),
                               "advantages": x["adv"][sl].to(dev), "ref_log_probs": x["ref"][sl].to(dev)},
                              meta_info={"temperature": 1.0})
        actor.update_policy(data)
        assert len(grabbed) == 1
        return grabbed[0]

    equal = 0.5 * (rank_dw((0, 8), 8, 0) + rank_dw((8, 16), 8, 0))
    uneven = 0.5 * (rank_dw((0, 10), 10, 8) + rank_dw((10, 16), 6, 8))
    assert rel(uneven, equal) < 2e-3
    want = O.fused_loss_reference(x["hidden"], x["weight"], x["labels"], x["old"], x["adv"], x["mask"], x["ref"],
                                  kl_penalty="low_var_kl", kl_coef=0.01, grad_accum=1.0, **CLIP)
    assert rel(uneven, want["dweight"]) < TOL_REL and rel(equal, want["dweight"]) < TOL_REL
