"""The oracle against (a) golden vectors produced by executing the reference (tests/golden/make_golden.py) and
(b) the live reference import when /root/reference is present. CPU only."""
import numpy as np
import pytest
import torch

from oracle import grpo_oracle as O

CLIP = (0.2, 0.3, 3.0)


def t(a):
    return torch.from_numpy(np.asarray(a))


# ------------------------------------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_advantage_golden(golden, tag):
    g = golden("advantage")
    uid = np.array([str(u) for u in g[f"{tag}_uid"]], dtype=object)
    adv, ret = O.compute_grpo_outcome_advantage(t(g[f"{tag}_rewards"]), t(g[f"{tag}_mask"]), uid)
    assert adv is ret
    np.testing.assert_array_equal(adv.numpy(), g[f"{tag}_adv"])  # same ops in the same order: bit-exact


def test_advantage_kat_values(golden):
    # SURVEY.md §8(c) KAT-A, typed in by hand: guards the golden file itself
    g = golden("advantage")
    want = [0.78334779, -0.63012475, -0.26111594, -0.21004158, 0.78334779, 1.47029090, -1.30557966, -0.63012475, 0, 0]
    np.testing.assert_allclose(g["a_adv"][:, 0], want, rtol=0, atol=1e-7)


def test_advantage_group_of_one_asserts():
    with pytest.raises(AssertionError):
        O.compute_grpo_outcome_advantage(torch.ones(3, 2), torch.ones(3, 2, dtype=torch.int64), np.array(["a", "a", "b"], dtype=object))


def test_policy_loss_golden(golden):
    g = golden("policy_loss")
    logp = t(g["a_logp"]).requires_grad_(True)
    res = O.compute_policy_loss(t(g["a_old"]), logp, t(g["a_adv"]), t(g["a_mask"]), *CLIP)
    res[0].backward()
    np.testing.assert_array_equal(np.array([float(r) for r in res], dtype=np.float32), g["a_out"])
    np.testing.assert_array_equal(logp.grad.numpy(), g["a_grad"])
    np.testing.assert_allclose(g["a_out"], [0.14340997, 0.28571430, 0.14285715, -0.18571429], atol=1e-7)  # KAT-B


def test_micro_batch_loss_golden(golden):
    g = golden("policy_loss")
    logp = t(g["b_logp"]).requires_grad_(True)
    loss, met = O.micro_batch_loss(logp, t(g["b_old"]), t(g["b_adv"]), t(g["b_mask"]), t(g["b_ref"]), grad_accum=4.0)
    loss.backward()
    got = [met["pg_only"], met["actor/pg_clipfrac_higher"], met["actor/pg_clipfrac_lower"], met["actor/ppo_kl"],
           met["actor/kl_loss"], met["actor/entropy_loss"], loss.detach()]
    np.testing.assert_array_equal(np.array([float(x) for x in got], dtype=np.float32), g["b_out"])
    np.testing.assert_array_equal(logp.grad.numpy(), g["b_grad"])


@pytest.mark.parametrize("mode", O.KL_MODES)
def test_compute_kl_golden(golden, mode):
    g = golden("policy_loss")
    lp = t(g["kl_logp"]).requires_grad_(True)
    v = O.compute_kl(lp, t(g["kl_ref"]), mode)
    v.sum().backward()
    np.testing.assert_array_equal(v.detach().numpy(), g[f"kl_{mode}"])
    np.testing.assert_array_equal(lp.grad.numpy(), g[f"kl_{mode}_grad"])


def test_kl_kat_values(golden):
    g = golden("policy_loss")  # SURVEY.md §8(c) KAT-C / D / E
    np.testing.assert_allclose(g["kl_low_var_kl"][:4], [0.01873076, 0.71828175, 0, 10.0], atol=1e-7)
    np.testing.assert_allclose(g["kl_low_var_kl_grad"][:4], [0.18126929, -1.71828175, 0, 0], atol=1e-6)
    assert float(g["masked_mean_zero"][0]) == 0.0
    np.testing.assert_allclose(g["clip_bounds"], [-0.2231435513, 0.2623642645], atol=1e-10)
    assert float(O.masked_mean(torch.ones(3, 4), torch.zeros(3, 4))) == 0.0
    with pytest.raises(NotImplementedError):
        O.compute_kl(torch.zeros(2), torch.zeros(2), "nope")


@pytest.mark.parametrize("tag", ["flat", "peaked"])
def test_end_to_end_golden(golden, tag):
    g = golden("end_to_end")
    uid = np.array([str(u) for u in g[f"{tag}_uid"]], dtype=object)
    adv, _ = O.compute_grpo_outcome_advantage(t(g[f"{tag}_rewards"]), t(g[f"{tag}_mask"]), uid)
    np.testing.assert_array_equal(adv.numpy(), g[f"{tag}_adv"])
    torch.set_num_threads(4)  # same reduction split as the generator
    res = O.fused_loss_reference(
        t(g[f"{tag}_hidden"]), t(g[f"{tag}_weight"]), t(g[f"{tag}_labels"]), t(g[f"{tag}_old"]), adv,
        t(g[f"{tag}_mask"]), t(g[f"{tag}_ref"]), temperature=float(g[f"{tag}_temp"][0]), grad_accum=2.0,
        want_entropy=True,
    )
    np.testing.assert_allclose(res["log_probs"].numpy(), g[f"{tag}_logp"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(res["entropy"].numpy(), g[f"{tag}_entropy"], rtol=0, atol=2e-5)
    m = res["metrics"]
    got = [m["actor/pg_loss"], m["actor/pg_clipfrac_higher"], m["actor/pg_clipfrac_lower"], m["actor/entropy_loss"],
           m["actor/ppo_kl"], m["actor/kl_loss"], m["pg_only"]]
    np.testing.assert_allclose([float(x) for x in got], g[f"{tag}_scalars"], rtol=1e-5, atol=1e-6)
    for key in ("dhidden", "dweight"):
        ref = g[f"{tag}_{key}"]
        rel = np.linalg.norm(res[key].numpy() - ref) / np.linalg.norm(ref)
        assert rel < 1e-5, (key, rel)


# ------------------------------------------------------------------------------------------------ live reference
def test_oracle_vs_live_reference(reference_modules):
    VF, ca = reference_modules
    g = torch.Generator().manual_seed(3)
    bsz, tl, n = 48, 17, 8
    r = O.synth_rollout(bsz, tl, 1000, n, seed=4, ragged=True)
    a_ref, _ = ca.compute_grpo_outcome_advantage(r["token_level_rewards"].clone(), r["response_mask"], r["uid"])
    a_ora, _ = O.compute_grpo_outcome_advantage(r["token_level_rewards"].clone(), r["response_mask"], r["uid"])
    assert torch.equal(a_ref, a_ora)
    z = torch.randn(bsz, tl, 333, generator=g) * 3
    lab = torch.randint(0, 333, (bsz, tl), generator=g)
    assert torch.equal(-VF.log_probs_from_logits(z, lab), O.log_probs_from_logits(z, lab))
    logp = O.log_probs_from_logits(z, lab)
    old = O.perturbed_log_probs(logp, seed=1, outlier_frac=0.05)
    ref = O.perturbed_log_probs(logp, seed=2, outlier_frac=0.05)
    for a, b in zip(ca.compute_policy_loss(old, logp, a_ref, r["response_mask"], *CLIP),
                    O.compute_policy_loss(old, logp, a_ref, r["response_mask"], *CLIP)):
        assert torch.equal(a, b)
    for mode in O.KL_MODES:
        assert torch.equal(ca.compute_kl(logp, ref, mode), O.compute_kl(logp, ref, mode))
    assert torch.equal(VF.masked_mean(logp, r["response_mask"]), O.masked_mean(logp, r["response_mask"]))
    assert torch.equal(VF.masked_mean(logp, r["response_mask"], dim=-1), O.masked_mean(logp, r["response_mask"], dim=-1))


# ------------------------------------------------------------------------------------------------ properties
def test_properties():
    h, w = O.synth_head(64, 64, 512, seed=9, sigma_w=0.3)
    lab = torch.randint(0, 512, (64,), generator=torch.Generator().manual_seed(0))
    logp, ent = O.lm_head_log_probs(h, w, lab, 1.0, want_entropy=True)
    assert (logp <= 0).all() and (ent >= 0).all() and (ent <= np.log(512) + 1e-5).all()
    adv = torch.randn(64)
    mask = torch.ones(64, dtype=torch.int64)
    # on-policy: ratio 1, nothing clipped, ppo_kl 0, d pg / d logp = -A / M
    lp = logp.clone().requires_grad_(True)
    pg, cfh, cfl, pkl = O.compute_policy_loss(logp.clone(), lp, adv, mask, *CLIP)
    pg.backward()
    assert float(cfh) == 0 and float(cfl) == 0 and float(pkl) == 0
    np.testing.assert_allclose(lp.grad.numpy(), (-adv / 64).numpy(), rtol=1e-6)
    # advantages are invariant under a row permutation; an all-equal group gives zeros
    r = O.synth_rollout(32, 9, 100, 8, seed=2, ragged=True)
    a0, _ = O.compute_grpo_outcome_advantage(r["token_level_rewards"].clone(), r["response_mask"], r["uid"])
    p = torch.randperm(32, generator=torch.Generator().manual_seed(1))
    a1, _ = O.compute_grpo_outcome_advantage(r["token_level_rewards"][p].clone(), r["response_mask"][p], r["uid"][p.numpy()])
    np.testing.assert_allclose(a1.numpy(), a0[p].numpy(), rtol=0, atol=1e-6)
    same = torch.zeros(4, 3)
    same[:, 2] = 0.5
    a2, _ = O.compute_grpo_outcome_advantage(same, torch.ones(4, 3, dtype=torch.int64), np.array(["g"] * 4, dtype=object))
    assert float(a2.abs().max()) == 0.0


def test_update_policy_reference_accumulates():
    bsz, tl, hd, v = 8, 6, 64, 256
    h, w = O.synth_head(bsz * tl, hd, v, seed=5)
    h = h.view(bsz, tl, hd)
    r = O.synth_rollout(bsz, tl, v, 4, seed=6, ragged=True)
    logp, _ = O.lm_head_log_probs(h, w, r["responses"])
    adv, _ = O.compute_grpo_outcome_advantage(r["token_level_rewards"].clone(), r["response_mask"], r["uid"])
    batch = {"responses": r["responses"], "response_mask": r["response_mask"], "advantages": adv,
             "old_log_probs": O.perturbed_log_probs(logp, seed=7), "ref_log_probs": O.perturbed_log_probs(logp, seed=8)}
    out = O.update_policy_reference(h, w, batch, global_batch_size_per_device=4, micro_batch_size_per_device_for_update=2)
    assert len(out["steps"]) == 2 and len(out["metrics"]["actor/pg_loss"]) == 4
    # the first optimizer step's dW equals the sum over its two micro-batches of (loss / GA).backward()
    acc = torch.zeros_like(w, dtype=torch.float32)
    for sl in (slice(0, 2), slice(2, 4)):
        res = O.fused_loss_reference(h[sl], w, batch["responses"][sl], batch["old_log_probs"][sl], adv[sl],
                                     batch["response_mask"][sl], batch["ref_log_probs"][sl], grad_accum=2.0)
        acc += res["dweight"]
    np.testing.assert_allclose(out["steps"][0]["dweight"].numpy(), acc.numpy(), rtol=1e-6, atol=1e-9)
