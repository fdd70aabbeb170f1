"""Parity at the HEADLINE head shape (Qwen2.5-VL-7B: H = 3584, V = 151 936 and the HF-padded 152 064) - the shape
BASELINE.json's metric is quoted on and the only one that takes the dW split-K tail plan (297 x 14 = 4158 tiles), the
2374-K-block dHidden tiles and, above 18 944 rows, a real chunk boundary.

Two checkers, both restatements of the reference arithmetic in fp32 (dp_actor.py:242-278, core_algos.py:291-353, 394-436,
torch_functional.py:45-71 through ``oracle/grpo_oracle.py``):

* the CPU oracle ``O.fused_loss_reference`` for 2048-row cases (seconds on the host cores);
* ``device_reference`` below for cases with more rows than one chunk: the SAME oracle functions evaluated on the GPU in
  fp32, the lm_head in row blocks (``F.linear`` in fp32, TF32 off) so that only ``block x V`` logits exist at a time.
  Test code may use torch matmul as its checker; the product may not and does not.

Tolerances are north_star's: log-probs / entropy 2e-3 absolute, loss and gradients 1e-2 relative.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import grpo_oracle as O

pytestmark = pytest.mark.gpu

H7B = 3584
V_QWEN, V_HF = 151936, 152064
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
    a, b = a.detach().float(), b.detach().float().to(a.device)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def device_reference(hidden, weight, labels, old, adv, mask, ref, *, temperature, kl_penalty, kl_coef, grad_accum,
                     want_entropy, block=2048):
    """fp32 restatement on the device, lm_head in row blocks. Pass 1: log-probs (+ entropy) of every row. Loss and
    dL/dlogp: the oracle's micro_batch_loss + autograd over the [rows] vector. Pass 2: each block's logits are rebuilt
    with autograd and dlogp is pushed through them into the hidden rows and an fp32 weight gradient."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        rows = hidden.shape[0]
        w32 = weight.float()
        logp = torch.empty(rows, device=hidden.device)
        ent = torch.empty(rows, device=hidden.device) if want_entropy else None
        with torch.no_grad():
            for r0 in range(0, rows, block):
                z = F.linear(hidden[r0:r0 + block].float(), w32) / temperature
                logp[r0:r0 + block] = O.log_probs_from_logits(z, labels[r0:r0 + block])
                if want_entropy:
                    ent[r0:r0 + block] = O.entropy_from_logits(z)
                del z
        lp = logp.clone().requires_grad_(True)
        loss, metrics = O.micro_batch_loss(lp, old, adv, mask, ref, kl_penalty=kl_penalty, kl_coef=kl_coef,
                                           grad_accum=grad_accum, **CLIP)
        loss.backward()
        dlogp = lp.grad
        wg = w32.clone().requires_grad_(True)
        dh = torch.empty(rows, hidden.shape[1], device=hidden.device)
        for r0 in range(0, rows, block):
            h = hidden[r0:r0 + block].float().requires_grad_(True)
            z = F.linear(h, wg) / temperature
            O.log_probs_from_logits(z, labels[r0:r0 + block]).backward(dlogp[r0:r0 + block])
            dh[r0:r0 + block] = h.grad
            del z, h
        return {"log_probs": logp, "entropy": ent, "loss": loss.detach(), "metrics": metrics, "dhidden": dh,
                "dweight": wg.grad}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def check_step(res, want, mask, *, use_ref, want_entropy):
    valid = mask.reshape(-1).bool()
    got_lp = res["log_probs"].reshape(-1)
    want_lp = want["log_probs"].reshape(-1).to(got_lp.device)
    assert float((got_lp - want_lp)[valid].abs().max()) < TOL_LOGP
    if want_entropy:
        got_e, want_e = res["entropy"].reshape(-1), want["entropy"].reshape(-1).to(got_lp.device)
        assert float((got_e - want_e)[valid].abs().max()) < TOL_LOGP
    m = res["metrics"].cpu()
    from spatialthinker_b200 import _lib

    wm = want["metrics"]
    assert abs(float(m[_lib.MET_SCALED]) - float(want["loss"])) <= TOL_REL * abs(float(want["loss"])) + 1e-7
    pairs = [(_lib.MET_TOTAL, "actor/pg_loss"), (_lib.MET_ENTROPY, "actor/entropy_loss"), (_lib.MET_PPO_KL, "actor/ppo_kl")]
    if use_ref:
        pairs.append((_lib.MET_KL_LOSS, "actor/kl_loss"))
    for slot, key in pairs:
        # entropy_loss and ppo_kl are masked MEANS OF LOG-PROBS (dp_actor.py:253, core_algos.py:352): their error is bounded
        # by the log-prob tolerance (2e-3 absolute), and ppo_kl is a near-cancelling mean (~5e-3 here). A tenth of that
        # bound is allowed on top of the 1e-2 relative one; the K = 3584 tensor-core accumulation alone shifts every
        # log-prob of a peaked row by ~5e-6 relative (measured 7e-5 absolute at log p ~ -15).
        slack = 0.1 * TOL_LOGP if key in ("actor/entropy_loss", "actor/ppo_kl") else 1e-5
        assert abs(float(m[slot]) - float(wm[key])) <= TOL_REL * abs(float(wm[key])) + slack, key
    for slot, key in ((_lib.MET_CLIPFRAC_HI, "actor/pg_clipfrac_higher"), (_lib.MET_CLIPFRAC_LO, "actor/pg_clipfrac_lower")):
        assert abs(float(m[slot]) - float(wm[key])) <= 2e-3, key  # a token exactly on a clip boundary may flip
    dh = res["dhidden"].reshape(-1, res["dhidden"].shape[-1])
    assert rel(dh, want["dhidden"].reshape(dh.shape)) < TOL_REL
    assert rel(res["dweight"], want["dweight"]) < TOL_REL
    if bool((~valid).any()):
        assert float(dh[~valid].abs().max()) == 0.0  # masked rows: exactly zero gradient


def make_case(bsz, tl, vocab, n, *, seed, sigma, ragged, dev, hdim=H7B):
    """Seeded synthetic micro-batch at the 7B head (SURVEY §8d generator): labels half sampled uniformly, half the row
    arg-max (so that both improbable and probable labels occur), old / ref = oracle log-probs + jitter + outliers."""
    hid, w = O.synth_head(bsz * tl, hdim, vocab, seed=seed, sigma_w=sigma)
    roll = O.synth_rollout(bsz, tl, vocab, n, seed=seed, ragged=ragged)
    adv, _ = O.compute_grpo_outcome_advantage(roll["token_level_rewards"].clone(), roll["response_mask"], roll["uid"])
    return {"hidden": hid.view(bsz, tl, hdim).to(dev), "weight": w.to(dev), "labels": roll["responses"].to(dev),
            "mask": roll["response_mask"].to(dev), "adv": adv.to(dev), "lens": roll["response_mask"].sum(-1)}


def perturb(logp, seed, dev):
    return O.perturbed_log_probs(logp.cpu(), seed=seed, outlier_frac=0.02).to(dev)


# ================================================================================================ (i) 2048 rows vs the CPU oracle
@pytest.mark.parametrize("vocab,sigma,temp", [(V_QWEN, 0.02, 1.0), (V_HF, 0.08, 0.9)])
def test_7b_head_2048_rows_vs_cpu_oracle(st, dev, vocab, sigma, temp):
    bsz, tl, n = 8, 256, 4
    x = make_case(bsz, tl, vocab, n, seed=vocab % 1000, sigma=sigma, ragged=True, dev=dev)
    hid_c, w_c, lab_c, mask_c, adv_c = (x[k].cpu() for k in ("hidden", "weight", "labels", "mask", "adv"))
    logp_c, _ = O.lm_head_log_probs(hid_c, w_c, lab_c, temp)
    old, ref = O.perturbed_log_probs(logp_c, seed=1, outlier_frac=0.02), O.perturbed_log_probs(logp_c, seed=2, outlier_frac=0.02)
    want = O.fused_loss_reference(hid_c, w_c, lab_c, old, adv_c, mask_c, ref, temperature=temp, kl_penalty="low_var_kl",
                                  kl_coef=1e-2, grad_accum=4.0, want_entropy=True, **CLIP)
    res = st.grpo_micro_batch_step(x["hidden"], x["weight"], x["labels"], old.to(dev), x["adv"], ref.to(dev), x["mask"],
                                   temperature=temp, kl_penalty="low_var_kl", kl_coef=1e-2, grad_accum=4.0,
                                   want_entropy=True, **CLIP)
    torch.cuda.synchronize()
    check_step(res, want, x["mask"], use_ref=True, want_entropy=True)
    # the forward-only entry (compute_log_prob path) on the same rows
    lp, ent = st.fused_lm_head_log_probs(x["hidden"], x["weight"], x["labels"], temp, want_entropy=True)
    assert float((lp.cpu() - want["log_probs"]).abs().max()) < TOL_LOGP
    assert float((ent.cpu() - want["entropy"]).abs().max()) < TOL_LOGP


def test_3b_head_config_c2_shape_vs_cpu_oracle(st, dev):
    """BASELINE configs[1]: the Qwen2.5-VL-3B head (H = 2048, V = 151 936) - 2048 rows of it against the CPU oracle. With
    K = 2048 the logits GEMM's tiles are 32 K-blocks long and the dHidden GEMM has 8 column blocks per row tile."""
    bsz, tl, n, hdim = 8, 256, 8, 2048
    x = make_case(bsz, tl, V_QWEN, n, seed=77, sigma=0.05, ragged=True, dev=dev, hdim=hdim)
    hid_c, w_c, lab_c, mask_c, adv_c = (x[k].cpu() for k in ("hidden", "weight", "labels", "mask", "adv"))
    logp_c, _ = O.lm_head_log_probs(hid_c, w_c, lab_c, 1.0)
    old, ref = O.perturbed_log_probs(logp_c, seed=1, outlier_frac=0.02), O.perturbed_log_probs(logp_c, seed=2, outlier_frac=0.02)
    want = O.fused_loss_reference(hid_c, w_c, lab_c, old, adv_c, mask_c, ref, temperature=1.0, kl_penalty="low_var_kl",
                                  kl_coef=1e-2, grad_accum=2.0, want_entropy=True, **CLIP)
    res = st.grpo_micro_batch_step(x["hidden"], x["weight"], x["labels"], old.to(dev), x["adv"], ref.to(dev), x["mask"],
                                   temperature=1.0, kl_penalty="low_var_kl", kl_coef=1e-2, grad_accum=2.0,
                                   want_entropy=True, **CLIP)
    torch.cuda.synchronize()
    check_step(res, want, x["mask"], use_ref=True, want_entropy=True)


# ================================================================================================ (ii) across a real chunk boundary
def test_7b_head_across_chunk_boundary_vs_device_reference(st, dev):
    """18 944 + 2 560 + 77 rows of the 7B head in ONE micro-batch: two chunks (the second ragged and not a multiple of
    the 512-row tile), the dW split-K tail, the progress windows, dW accumulated across the chunk boundary."""
    from spatialthinker_b200 import _lib

    cap = int(_lib.load().grpo_chunk_capacity_rows())
    rows = cap + 2560 + 77
    g = torch.Generator().manual_seed(2024)
    hid = torch.randn(rows, H7B, generator=g).to(torch.bfloat16).to(dev)
    w = (0.03 * torch.randn(V_QWEN, H7B, generator=g)).to(torch.bfloat16).to(dev)
    lab = torch.randint(0, V_QWEN, (rows,), generator=g).to(dev)
    mask = (torch.rand(rows, generator=g) > 0.25).long().to(dev)
    adv = torch.randn(rows, generator=g).to(dev)
    lp0, _ = st.fused_lm_head_log_probs(hid, w, lab, 1.0)
    old, ref = perturb(lp0, 11, dev), perturb(lp0, 12, dev)
    kw = dict(temperature=1.0, kl_penalty="low_var_kl", kl_coef=1e-2, grad_accum=2.0)
    want = device_reference(hid, w, lab, old, adv, mask, ref, want_entropy=False, **kw)
    res = st.grpo_micro_batch_step(hid, w, lab, old, adv, ref, mask, **kw, **CLIP)
    torch.cuda.synchronize()
    check_step(res, want, mask, use_ref=True, want_entropy=False)
    # each chunk on its own
    dh = res["dhidden"]
    for sl in (slice(0, cap), slice(cap, rows)):
        assert rel(dh[sl], want["dhidden"][sl]) < TOL_REL


# ================================================================================================ (iii) C4 / C5 flavours
def test_7b_head_c4_flavour(st, dev):
    """Config C4: 2048-token responses, reference-policy low_var_kl and the true per-token entropy output - 10 sequences
    = 20 480 rows (> one chunk)."""
    bsz, tl, n = 10, 2048, 2
    x = make_case(bsz, tl, V_QWEN, n, seed=44, sigma=0.03, ragged=False, dev=dev)
    lp0, _ = st.fused_lm_head_log_probs(x["hidden"], x["weight"], x["labels"], 1.0)
    old, ref = perturb(lp0, 21, dev), perturb(lp0, 22, dev)
    kw = dict(temperature=1.0, kl_penalty="low_var_kl", kl_coef=1e-2, grad_accum=8.0)
    flat = lambda t_: t_.reshape(-1, *t_.shape[2:])  # noqa: E731
    want = device_reference(flat(x["hidden"]), x["weight"], flat(x["labels"]), flat(old), flat(x["adv"]), flat(x["mask"]),
                            flat(ref), want_entropy=True, **kw)
    res = st.grpo_micro_batch_step(x["hidden"], x["weight"], x["labels"], old, x["adv"], ref, x["mask"], want_entropy=True,
                                   **kw, **CLIP)
    torch.cuda.synchronize()
    check_step(res, want, x["mask"], use_ref=True, want_entropy=True)
    from spatialthinker_b200 import _lib

    m = res["metrics"].cpu()
    want_h = float((want["entropy"] * flat(x["mask"])).sum() / flat(x["mask"]).sum())
    assert abs(float(m[_lib.MET_TRUE_ENTROPY]) - want_h) <= TOL_LOGP
    assert 0.0 < want_h <= math.log(V_QWEN)


def test_7b_head_c5_flavour(st, dev):
    """Config C5: 4096-token ragged responses, groups of 16, padded slots compacted away on the device
    (``valid_rows``) - 16 sequences = 65 536 slots, about half of them valid (two chunks after compaction)."""
    bsz, tl, n = 16, 4096, 16
    x = make_case(bsz, tl, V_QWEN, n, seed=55, sigma=0.02, ragged=True, dev=dev)
    valid_rows = int(x["lens"].sum())
    assert 18944 < valid_rows < bsz * tl
    lp0, _ = st.fused_lm_head_log_probs(x["hidden"], x["weight"], x["labels"], 1.0)
    old, ref = perturb(lp0, 31, dev), perturb(lp0, 32, dev)
    kw = dict(temperature=1.0, kl_penalty="low_var_kl", kl_coef=1e-2, grad_accum=4.0)
    flat = lambda t_: t_.reshape(-1, *t_.shape[2:])  # noqa: E731
    want = device_reference(flat(x["hidden"]), x["weight"], flat(x["labels"]), flat(old), flat(x["adv"]), flat(x["mask"]),
                            flat(ref), want_entropy=False, **kw)
    res = st.grpo_micro_batch_step(x["hidden"], x["weight"], x["labels"], old, x["adv"], ref, x["mask"],
                                   valid_rows=valid_rows, **kw, **CLIP)
    torch.cuda.synchronize()
    check_step(res, want, x["mask"], use_ref=True, want_entropy=False)
    # compaction must not change anything but the padded slots (zeros there): same call without the hint
    dense = st.grpo_micro_batch_step(x["hidden"], x["weight"], x["labels"], old, x["adv"], ref, x["mask"], **kw, **CLIP)
    valid = x["mask"].bool()
    assert float((dense["log_probs"] - res["log_probs"])[valid].abs().max()) < 1e-5
    assert rel(dense["dweight"], res["dweight"]) < 2e-3
    assert rel(dense["dhidden"], res["dhidden"]) < 2e-3
    # group statistics of the n = 16 groups on the device against the oracle values used above
    roll = O.synth_rollout(bsz, tl, V_QWEN, n, seed=55, ragged=True)
    got, _ = st.compute_grpo_outcome_advantage(roll["token_level_rewards"].to(dev), roll["response_mask"].to(dev), roll["uid"])
    assert bool(((got - x["adv"]).abs() <= 1e-6 * x["adv"].abs().clamp_min(1.0)).all())


# ================================================================================================ the device reference itself
def test_device_reference_matches_cpu_oracle(dev):
    """The row-blocked fp32 device restatement used above is itself pinned to the CPU oracle on a small case."""
    bsz, tl, h, v = 4, 48, 128, 4096
    hid, w = O.synth_head(bsz * tl, h, v, seed=3, sigma_w=0.1)
    roll = O.synth_rollout(bsz, tl, v, 2, seed=3, ragged=True)
    lab, mask = roll["responses"].reshape(-1), roll["response_mask"].reshape(-1)
    logp, _ = O.lm_head_log_probs(hid, w, lab)
    old, ref = O.perturbed_log_probs(logp, seed=1, outlier_frac=0.05), O.perturbed_log_probs(logp, seed=2, outlier_frac=0.05)
    adv = torch.randn(bsz * tl, generator=torch.Generator().manual_seed(1))
    kw = dict(temperature=0.9, kl_penalty="low_var_kl", kl_coef=0.02, grad_accum=2.0)
    want = O.fused_loss_reference(hid, w, lab, old, adv, mask, ref, want_entropy=True, **kw, **CLIP)
    got = device_reference(hid.to(dev), w.to(dev), lab.to(dev), old.to(dev), adv.to(dev), mask.to(dev), ref.to(dev),
                           want_entropy=True, block=50, **kw)
    np.testing.assert_allclose(got["log_probs"].cpu().numpy(), want["log_probs"].numpy(), atol=2e-5)
    np.testing.assert_allclose(got["entropy"].cpu().numpy(), want["entropy"].numpy(), atol=2e-5)
    assert abs(float(got["loss"]) - float(want["loss"])) <= 1e-5 * abs(float(want["loss"]))
    assert rel(got["dhidden"], want["dhidden"]) < 1e-4 and rel(got["dweight"], want["dweight"]) < 1e-4
