"""Parity of the CUDA path (through the Python mirror of the reference interface and the C ABI underneath) against
the CPU oracle and the golden vectors generated from the reference.

Tolerances are the ones BASELINE.json's north_star states: per-token log-probs and entropy 2e-3 absolute,
advantages 1e-6 (|a - a_ref| <= 1e-6 * max(1, |a_ref|)), loss and gradients 1e-2 relative (Frobenius for tensors).
"""
import ctypes
import math
import os

import numpy as np
import pytest
import torch

from oracle import grpo_oracle as O

pytestmark = pytest.mark.gpu

CLIP = (0.2, 0.3, 3.0)
EPI_SHARE_DEFAULT = 0  # library default of the "epi_share" option (restored by tests that toggle it)
DH_SPLIT_DEFAULT = int(os.environ.get("GRPO_DH_SPLIT", "1") != "0")  # library default of "dh_split" (env override as in the library)
TOL_LOGP = 2e-3
TOL_ADV = 1e-6
TOL_REL = 1e-2


@pytest.fixture(scope="module")
def st():
    import spatialthinker_b200 as st

    st.load_library()
    return st


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def t(x):
    return torch.from_numpy(np.asarray(x))


def adv_close(got, want):
    got, want = got.detach().cpu(), want.detach().cpu()
    assert bool(((got - want).abs() <= TOL_ADV * want.abs().clamp_min(1.0)).all()), float((got - want).abs().max())


# ================================================================================================ tcgen05 GEMM
@pytest.mark.parametrize("cta", [1, 2])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0), (2, 1), (3, 1), (2, 0), (3, 0)])
def test_debug_gemm_layouts(st, dev, cta, a_mn, b_mn):
    from spatialthinker_b200 import _lib

    lib = _lib.load()
    m, n, k = 648, 776, 456  # ragged in every dimension (multiples of 8 for the 16-byte TMA pitch)
    g = torch.Generator().manual_seed(m + 7 * cta)
    a = torch.randn(m, k, generator=g).to(torch.bfloat16)
    b = torch.randn(n, k, generator=g).to(torch.bfloat16)
    want = a.double() @ b.double().t()
    def blocked(x):  # [r][c] -> [r/64][c/64][64][64], zero padded
        r, c = x.shape
        rp, cp = -(-r // 64) * 64, -(-c // 64) * 64
        xp = torch.zeros(rp, cp, dtype=x.dtype)
        xp[:r, :c] = x
        return xp.view(rp // 64, 64, cp // 64, 64).permute(0, 2, 1, 3).contiguous()

    a_d = {0: a, 1: a.t().contiguous(), 2: blocked(a), 3: blocked(a.t())}[a_mn].to(dev)
    b_d = (b.t().contiguous() if b_mn else b).to(dev)
    c = torch.full((m, n), 3.0, device=dev)
    _lib.check(lib.grpo_debug_gemm(a_d.data_ptr(), b_d.data_ptr(), c.data_ptr(), m, n, k, a_mn, b_mn, cta, 1,
                                   _lib.stream_ptr(dev)), "gemm")
    torch.cuda.synchronize()
    np.testing.assert_allclose(c.cpu().double().numpy(), (want + 3.0).numpy(), rtol=0, atol=2e-3)


@pytest.mark.parametrize("accumulate", [0, 1])
@pytest.mark.parametrize("tma", [0, 1])
def test_debug_gemm_fp32_epilogues(st, dev, tma, accumulate):
    """The fp32 epilogue through shared memory + bulk tensor store / reduce-add (the dW GEMM's) against the per-thread
    st.global / red.global one, ragged edges in both dimensions (the TMA unit clips the boxes)."""
    from spatialthinker_b200 import _lib

    lib = _lib.load()
    m, n, k = 1000, 520, 200
    g = torch.Generator().manual_seed(11)
    a = torch.randn(m, k, generator=g).to(torch.bfloat16)
    b = torch.randn(n, k, generator=g).to(torch.bfloat16)
    want = a.double() @ b.double().t() + (2.5 if accumulate else 0.0)
    c = torch.full((m + 8, n), 2.5, device=dev)  # 8 guard rows: nothing may be written past row m
    a_d, b_d = a.to(dev), b.to(dev)
    _lib.check(lib.grpo_set_option(b"dw_tma", tma), "set_option")
    try:
        _lib.check(lib.grpo_debug_gemm(a_d.data_ptr(), b_d.data_ptr(), c.data_ptr(), m, n, k, 0, 0, 2,
                                       accumulate, _lib.stream_ptr(dev)), "gemm")
        torch.cuda.synchronize()
    finally:
        lib.grpo_set_option(b"dw_tma", 1)
    np.testing.assert_allclose(c[:m].cpu().double().numpy(), want.numpy(), rtol=0, atol=2e-3)
    assert bool((c[m:] == 2.5).all())


@pytest.mark.parametrize("cta,ksub_rows", [(2, 512), (1, 128)])
@pytest.mark.parametrize("a_mode,b_mn", [(0, 0), (3, 1)])
def test_debug_gemm_split_k_tail(st, dev, cta, ksub_rows, a_mode, b_mn):
    """More tiles than persistent CTA groups with a small remainder: the last round is cut along K into slices that the
    accumulating epilogue adds up (the dW GEMM's tail). Same result as without the split, and as fp64 torch."""
    from spatialthinker_b200 import _lib

    lib = _lib.load()
    groups = 148 // cta
    tiles_n = 3
    tiles_m = (groups + 5 + tiles_n - 1) // tiles_n + 1  # a handful of tiles beyond one full round
    m, n, k = tiles_m * ksub_rows - 37, tiles_n * 256 - 8, 64 * 23 + 40  # ragged in every dimension, 24 K-blocks
    g = torch.Generator().manual_seed(17)
    a = torch.randn(m, k, generator=g).to(torch.bfloat16)
    b = torch.randn(n, k, generator=g).to(torch.bfloat16)
    want = a.double() @ b.double().t() + 1.0
    if a_mode == 3:  # blocked [k/64][m/64][64 k][64 m] image of A, as the dW GEMM reads the stash
        kp, mp = (k + 63) // 64 * 64, (m + 63) // 64 * 64
        pad = torch.zeros(mp, kp, dtype=torch.bfloat16)
        pad[:m, :k] = a
        a_d = pad.view(mp // 64, 64, kp // 64, 64).permute(2, 0, 3, 1).contiguous().to(dev)
    else:
        a_d = a.to(dev)
    b_d = (b.t().contiguous() if b_mn else b).to(dev)
    outs = []
    for split in (1, 0):
        c = torch.full((m, n), 1.0, device=dev)
        _lib.check(lib.grpo_set_option(b"dw_split", split), "set_option")
        try:
            _lib.check(lib.grpo_debug_gemm(a_d.data_ptr(), b_d.data_ptr(), c.data_ptr(), m, n, k, a_mode, b_mn, cta, 1,
                                           _lib.stream_ptr(dev)), "gemm")
            torch.cuda.synchronize()
        finally:
            lib.grpo_set_option(b"dw_split", 1)
        outs.append(c.cpu())
    np.testing.assert_allclose(outs[0].double().numpy(), want.numpy(), rtol=0, atol=5e-3)
    np.testing.assert_allclose(outs[0].numpy(), outs[1].numpy(), rtol=0, atol=2e-3)  # fp32 summation order differs


@pytest.mark.parametrize("shape", ["few_tiles", "big_remainder"])
@pytest.mark.parametrize("a_mode,b_mn,tma", [(2, 1, 1), (0, 0, 1), (2, 1, 0)])
def test_debug_gemm_multi_round_split(st, dev, shape, a_mode, b_mn, tma):
    """Split-K plan of the dHidden GEMM's fp32 path (option dw_split = 2 on the debug entry): the tiles of the last partial
    round - or ALL tiles when there are fewer tiles than CTA pairs - are cut along K into slices walked slice-major over
    several short rounds; the accumulating epilogue adds the slices up. (2, 1) = blocked K-major A, B read transposed:
    the operand layouts of the dHidden GEMM."""
    from spatialthinker_b200 import _lib

    lib = _lib.load()
    if shape == "few_tiles":  # 3 x 2 wide tiles on 74 pairs, 24 K-blocks (ragged everywhere)
        m, n, k = 3 * 512 - 37, 2 * 256 - 8, 64 * 23 + 40
    else:  # 40 x 3 = 120 tiles on 74 pairs: remainder 46 > half a round, 64 K-blocks
        m, n, k = 40 * 512 - 100, 3 * 256 - 16, 64 * 63 + 24
    info = (ctypes.c_int32 * 4)()
    _lib.check(lib.grpo_debug_plan_units(-(-m // 512) * -(-n // 256), -(-k // 64), 74, 2, None, 0, info), "plan")
    assert info[3] > 1, "this shape is meant to take the split plan"
    g = torch.Generator().manual_seed(23)
    a = torch.randn(m, k, generator=g).to(torch.bfloat16)
    b = torch.randn(n, k, generator=g).to(torch.bfloat16)
    want = a.double() @ b.double().t() + 1.0
    if a_mode == 2:  # blocked [m/64][k/64][64 m][64 k] image of A, as the dHidden GEMM reads the stash
        kp, mp = (k + 63) // 64 * 64, (m + 63) // 64 * 64
        pad = torch.zeros(mp, kp, dtype=torch.bfloat16)
        pad[:m, :k] = a
        a_d = pad.view(mp // 64, 64, kp // 64, 64).permute(0, 2, 1, 3).contiguous().to(dev)
    else:
        a_d = a.to(dev)
    b_d = (b.t().contiguous() if b_mn else b).to(dev)
    outs = []
    for split in (2, 0):
        c = torch.full((m + 8, n), 1.0, device=dev)  # 8 guard rows
        _lib.check(lib.grpo_set_option(b"dw_split", split), "set_option")
        _lib.check(lib.grpo_set_option(b"dw_tma", tma), "set_option")
        try:
            _lib.check(lib.grpo_debug_gemm(a_d.data_ptr(), b_d.data_ptr(), c.data_ptr(), m, n, k, a_mode, b_mn, 2, 1,
                                           _lib.stream_ptr(dev)), "gemm")
            torch.cuda.synchronize()
        finally:
            lib.grpo_set_option(b"dw_split", 1)
            lib.grpo_set_option(b"dw_tma", 1)
        assert bool((c[m:] == 1.0).all())
        outs.append(c[:m].cpu())
    np.testing.assert_allclose(outs[0].double().numpy(), want.numpy(), rtol=0, atol=5e-3)
    np.testing.assert_allclose(outs[0].numpy(), outs[1].numpy(), rtol=0, atol=2e-3)  # fp32 summation order differs


@pytest.mark.parametrize("rows,h,v", [(4096 - 33, 3584, 32768 + 72), (1100, 256, 2 * 4096 + 520), (9000, 512, 4096)])
@pytest.mark.parametrize("dent", [False, True])
def test_dhidden_split_path_matches_direct(st, dev, rows, h, v, dent):
    """Option dh_split: when the dHidden GEMM's tiles are not a whole number of rounds over the CTA pairs it accumulates in
    fp32 with a split-K tail and converts in a fix-up pass. Same gradients as the direct bf16 epilogue (bf16 rounding
    of sums taken in a different order) and as the oracle. First shape = the reference's 4-sequence micro-batch at the
    7B head width (8 x 14 tiles = 1.51 rounds), second = fewer tiles than pairs, third = 18 x 2 tiles."""
    from spatialthinker_b200 import _lib

    lib = _lib.load()
    hid, w = O.synth_head(rows, h, v, seed=5, sigma_w=0.1)
    g = torch.Generator().manual_seed(5)
    lab = torch.randint(0, v, (rows,), generator=g)
    lab[7] = -100  # a label no vocabulary row matches
    gl = torch.randn(rows, generator=g) / rows
    gl[::5] = 0.0  # masked rows
    ge = torch.randn(rows, generator=g) / rows if dent else None
    outs, launches = [], []
    try:
        for split in (1, 0):
            _lib.check(lib.grpo_set_option(b"dh_split", split), "set_option")
            n0 = lib.grpo_launch_count()
            hd, wd = hid.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True)
            lp, ent = st.fused_lm_head_log_probs(hd, wd, lab.clamp_min(0).to(dev) if dent else lab.to(dev), 1.3,
                                                 want_entropy=dent)
            loss = (lp * gl.to(dev)).sum()
            if dent:
                loss = loss + (ent * ge.to(dev)).sum()
            loss.backward()
            torch.cuda.synchronize()
            outs.append((hd.grad.clone(), wd.grad.clone()))
            launches.append(lib.grpo_launch_count() - n0)
    finally:
        lib.grpo_set_option(b"dh_split", DH_SPLIT_DEFAULT)
    assert launches[0] == launches[1] + 1, "the split path adds exactly the fix-up kernel (one chunk)"
    assert rel(outs[0][0], outs[1][0]) < 2e-3  # bf16 outputs, fp32 sums in another order
    assert rel(outs[0][1], outs[1][1]) < 1e-5  # dW does not go through the split path
    if not dent:
        valid = lab >= 0
        hf, wf = hid.float().requires_grad_(True), w.float().requires_grad_(True)
        lp_ref, _ = O.lm_head_log_probs(hf, wf, lab.clamp_min(0), 1.3)
        (lp_ref * gl * valid).sum().backward()
        keep = valid.nonzero().squeeze(1)
        assert rel(outs[0][0][keep.to(dev)], hf.grad[keep]) < TOL_REL


@pytest.mark.parametrize("rows,h,v", [(1500, 2560, 8192 + 136), (4608, 2304, 4096)])
def test_serpentine_k_order_matches_forward_order(st, dev, rows, h, v):
    """Option k_serp: every other whole tile of a CTA pair walks its K-blocks backwards (dW GEMM: bit 0, dHidden GEMM:
    bit 1) so that the operand every round re-reads is met again where it was touched last. Same sums in another order:
    dW and dHidden (both handed back as bf16) agree to one bf16 rounding. First shape: 17 x 10 dW tiles = 2.3 rounds
    over the 74 CTA pairs (dHidden: one round, untouched); second: 9 x 9 dHidden tiles = 1.1 rounds on the direct bf16
    path (dW: one round, untouched)."""
    from spatialthinker_b200 import _lib

    lib = _lib.load()
    hid, w = O.synth_head(rows, h, v, seed=9, sigma_w=0.1)
    g = torch.Generator().manual_seed(9)
    lab = torch.randint(0, v, (rows,), generator=g)
    gl = torch.randn(rows, generator=g) / rows
    outs = []
    try:
        _lib.check(lib.grpo_set_option(b"dh_split", 0), "set_option")
        for serp in (0, 3):
            _lib.check(lib.grpo_set_option(b"k_serp", serp), "set_option")
            hd, wd = hid.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True)
            lp, _ = st.fused_lm_head_log_probs(hd, wd, lab.to(dev), 1.0)
            (lp * gl.to(dev)).sum().backward()
            torch.cuda.synchronize()
            outs.append((lp.detach().clone(), hd.grad.clone(), wd.grad.clone()))
    finally:
        lib.grpo_set_option(b"k_serp", 0)
        lib.grpo_set_option(b"dh_split", DH_SPLIT_DEFAULT)
    assert torch.equal(outs[0][0], outs[1][0])     # the logits GEMM is untouched
    assert rel(outs[0][1], outs[1][1]) < 2e-3      # bf16 outputs of fp32 sums taken in another order
    assert rel(outs[0][2], outs[1][2]) < 2e-3      # autograd hands dW back in the parameter's dtype (bf16)
    pairs = 74
    if -(-v // 512) * -(-h // 256) > pairs:        # more than one round of dW tiles: the order of some sums changed
        assert not torch.equal(outs[0][2], outs[1][2]), "the option did not reach the dW GEMM"
    if -(-rows // 512) * -(-h // 256) > pairs:
        assert not torch.equal(outs[0][1], outs[1][1]), "the option did not reach the dHidden GEMM"


def test_epilogue_variants_agree(st, dev):
    """Softmax-epilogue variants (plain loop / pipelined TMEM drain / stash through bulk tensor stores) and dW-epilogue
    variants must give the same log-probs bit for bit and the same gradients up to fp32 accumulation order."""
    from spatialthinker_b200 import _lib

    lib = _lib.load()
    rows, h, v = 1100, 256, 2 * 4096 + 520  # several full vocab tiles + a ragged last one; rows ragged too
    hid, w = O.synth_head(rows, h, v, seed=3, sigma_w=0.1)
    g = torch.Generator().manual_seed(3)
    lab = torch.randint(0, v, (rows,), generator=g)
    gl = torch.randn(rows, generator=g) / rows
    hf, wf = hid.float().requires_grad_(True), w.float().requires_grad_(True)
    lp_ref, _ = O.lm_head_log_probs(hf, wf, lab, 1.0)
    (lp_ref * gl).sum().backward()
    outs = {}
    try:
        # dHidden is compared bit for bit below: keep it on its one-writer-per-tile epilogue (the split-K path sums the
        # K slices of a tile in whatever order they finish; it has its own test)
        _lib.check(lib.grpo_set_option(b"dh_split", 0), "set_option")
        # (epi_mode, dw_tma, acc_lead, epi_share); epi_share = both epilogue warpgroups on one accumulator at a time
        for epi, dwt, lead, share in ((0, 0, 0, 0), (1, 0, 0, 0), (3, 0, 1, 0), (3, 1, 2, 0), (0, 1, 3, 0), (7, 1, 2, 0),
                                      (0, 1, 2, 1), (1, 1, 0, 1), (3, 1, 2, 1), (7, 1, 2, 1), (7, 1, 3, 1)):
            _lib.check(lib.grpo_set_option(b"epi_mode", epi), "set_option")
            _lib.check(lib.grpo_set_option(b"dw_tma", dwt), "set_option")
            _lib.check(lib.grpo_set_option(b"acc_lead", lead), "set_option")
            _lib.check(lib.grpo_set_option(b"epi_share", share), "set_option")
            hd, wd = hid.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True)
            lp, ent = st.fused_lm_head_log_probs(hd, wd, lab.to(dev), 1.0, want_entropy=True)
            (lp * gl.to(dev)).sum().backward()
            outs[(epi, dwt, lead, share)] = (lp.detach().clone(), hd.grad.clone(), wd.grad.clone(), ent.detach().clone())
    finally:
        lib.grpo_set_option(b"epi_mode", 7)
        lib.grpo_set_option(b"dw_tma", 1)
        lib.grpo_set_option(b"acc_lead", 2)
        lib.grpo_set_option(b"epi_share", EPI_SHARE_DEFAULT)
        lib.grpo_set_option(b"dh_split", DH_SPLIT_DEFAULT)
    base = outs[(0, 0, 0, 0)]
    assert float((base[0].cpu() - lp_ref.detach()).abs().max()) < TOL_LOGP
    assert rel(base[1], hf.grad) < TOL_REL and rel(base[2], wf.grad) < TOL_REL
    for key, (lp, dh, dw, ent) in outs.items():
        if key[3] == 0:
            assert torch.equal(lp, base[0]), key
            assert torch.equal(dh, base[1]), key  # same stash bits -> same dHidden bits
        else:  # the row sums are formed from two half-tile partials: same values up to fp32 summation order
            assert float((lp - base[0]).abs().max()) < 1e-5, key
            assert float((ent - base[3]).abs().max()) < 1e-4, key
            assert rel(dh, base[1]) < 1e-3, key
        assert rel(dw, base[2]) < 1e-3, key  # bf16 grads: an fp32 ulp may flip a rounding


# ================================================================================================ logits surface
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_log_probs_from_logits(st, dev, dtype):
    g = torch.Generator().manual_seed(0)
    z = (3 * torch.randn(5, 37, 1000, generator=g)).to(dtype)
    lab = torch.randint(0, 1000, (5, 37), generator=g)
    want = O.log_probs_from_logits(z.float(), lab)
    zd = z.to(dev).requires_grad_(True)
    got = st.log_probs_from_logits(zd, lab.to(dev))
    assert got.shape == (5, 37) and got.dtype == torch.float32 and bool((got <= 0).all())
    assert float((got.cpu() - want).abs().max()) < 1e-4
    assert st.logprobs_from_logits is st.log_probs_from_logits
    # backward: d/dz sum(w * logp)
    w = torch.randn(5, 37, generator=g)
    (got * w.to(dev)).sum().backward()
    zr = z.float().requires_grad_(True)
    (O.log_probs_from_logits(zr, lab) * w).sum().backward()
    assert rel(zd.grad, zr.grad) < (1e-5 if dtype == torch.float32 else 6e-3)
    assert float(zd.grad.float().sum(-1).abs().max()) < 2e-2  # sum_v dL/dz = 0


def test_entropy_from_logits(st, dev):
    g = torch.Generator().manual_seed(1)
    z = 4 * torch.randn(64, 151936 // 8, generator=g)
    zd = z.to(dev).requires_grad_(True)
    ent = st.entropy_from_logits(zd)
    want = O.entropy_from_logits(z)
    assert float((ent.cpu() - want).abs().max()) < 1e-4
    assert bool((ent >= 0).all()) and bool((ent <= math.log(z.shape[-1]) + 1e-4).all())
    ent.sum().backward()
    zr = z.clone().requires_grad_(True)
    O.entropy_from_logits(zr).sum().backward()
    assert rel(zd.grad, zr.grad) < 1e-4


def test_logits_edge_cases(st, dev):
    # odd vocab (scalar path), unaligned rows, single row, empty batch
    g = torch.Generator().manual_seed(2)
    z = torch.randn(3, 1001, generator=g)
    lab = torch.tensor([0, 1000, 500])
    got = st.log_probs_from_logits(z.to(dev), lab.to(dev))
    assert float((got.cpu() - O.log_probs_from_logits(z, lab)).abs().max()) < 1e-5
    empty = st.log_probs_from_logits(torch.zeros(0, 16, device=dev), torch.zeros(0, dtype=torch.int64, device=dev))
    assert empty.shape == (0,)
    with pytest.raises(ValueError):
        st.log_probs_from_logits(torch.zeros(2, 16, device=dev), torch.zeros(3, dtype=torch.int64, device=dev))


def test_masked_mean(st, dev, golden):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(7, 300, generator=g)
    for mask in ((torch.rand(7, 300, generator=g) > 0.4).long(), (torch.rand(7, 300, generator=g) > 0.4).float(),
                 torch.rand(7, 300, generator=g) > 0.4):
        got = st.masked_mean(x.to(dev), mask.to(dev))
        want = O.masked_mean(x, mask.float() if mask.dtype == torch.bool else mask)
        assert abs(float(got) - float(want)) < 1e-6
        np.testing.assert_allclose(st.masked_mean(x.to(dev), mask.to(dev), dim=-1).cpu().numpy(),
                                   O.masked_mean(x, mask.float(), dim=-1).numpy(), rtol=1e-5, atol=1e-6)
    assert float(st.masked_mean(torch.ones(3, 4, device=dev), torch.zeros(3, 4, device=dev))) == 0.0  # KAT-D
    xd = x.to(dev).requires_grad_(True)
    m = (torch.rand(7, 300, generator=g) > 0.5).long()
    st.masked_mean(xd, m.to(dev)).backward()
    np.testing.assert_allclose(xd.grad.cpu().numpy(), (m.float() / m.sum()).numpy(), rtol=1e-6)


# ================================================================================================ core_algos surface
@pytest.mark.parametrize("mode", O.KL_MODES)
def test_compute_kl_golden(st, dev, golden, mode):
    g = golden("policy_loss")
    lp = t(g["kl_logp"]).to(dev).requires_grad_(True)
    out = st.compute_kl(lp, t(g["kl_ref"]).to(dev), mode)
    out.sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), g[f"kl_{mode}"], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(lp.grad.cpu().numpy(), g[f"kl_{mode}_grad"], rtol=2e-6, atol=1e-6)
    assert st.kl_penalty(lp.detach(), t(g["kl_ref"]).to(dev), mode).shape == lp.shape


def test_policy_loss_golden(st, dev, golden):
    g = golden("policy_loss")
    lp = t(g["a_logp"]).to(dev).requires_grad_(True)
    res = st.compute_policy_loss(t(g["a_old"]).to(dev), lp, t(g["a_adv"]).to(dev), t(g["a_mask"]).to(dev), *CLIP)
    res[0].backward()
    np.testing.assert_allclose([float(r) for r in res], g["a_out"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(lp.grad.cpu().numpy(), g["a_grad"], rtol=1e-6, atol=1e-8)


def test_policy_loss_seeded_vs_oracle(st, dev):
    g = torch.Generator().manual_seed(4)
    bsz, tl = 16, 700
    lp = -3 * torch.rand(bsz, tl, generator=g)
    old = O.perturbed_log_probs(lp, seed=1, jitter=0.2, outlier_frac=0.05)
    adv = torch.randn(bsz, 1, generator=g).expand(bsz, tl).contiguous()
    lens = torch.randint(1, tl + 1, (bsz,), generator=g)
    mask = (torch.arange(tl)[None] < lens[:, None]).long()
    lpr = lp.clone().requires_grad_(True)
    want = O.compute_policy_loss(old, lpr, adv, mask, *CLIP)
    (want[0] + 0.3 * want[3]).backward()
    lpd = lp.to(dev).requires_grad_(True)
    got = st.compute_policy_loss(old.to(dev), lpd, adv.to(dev), mask.to(dev), *CLIP)
    (got[0] + 0.3 * got[3]).backward()
    for a, b in zip(got, want):
        assert abs(float(a) - float(b)) <= 1e-5 * max(1.0, abs(float(b)))
    assert float(want[1]) > 0 and float(want[2]) > 0  # both clip sides and the dual clip are exercised
    assert rel(lpd.grad, lpr.grad) < 1e-5
    # on-policy: ratio 1 => nothing clipped, ppo_kl 0
    res = st.compute_policy_loss(lp.to(dev), lp.to(dev), adv.to(dev), mask.to(dev), *CLIP)
    assert float(res[1]) == 0 and float(res[2]) == 0 and float(res[3]) == 0


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_advantage_golden(st, dev, golden, tag):
    g = golden("advantage")
    uid = np.array([str(u) for u in g[f"{tag}_uid"]], dtype=object)
    adv, ret = st.compute_grpo_outcome_advantage(t(g[f"{tag}_rewards"]).to(dev), t(g[f"{tag}_mask"]).to(dev), uid)
    assert adv is ret and adv.dtype == torch.float32
    adv_close(adv, t(g[f"{tag}_adv"]))


def test_advantage_full_size_vs_oracle(st, dev):
    # config C3's batch: 512 prompts x n=8 = 4096 sequences, T=1024, permuted groups, ragged masks
    roll = O.synth_rollout(4096, 1024, 151936, 8, seed=9, ragged=True)
    want, _ = O.compute_grpo_outcome_advantage(roll["token_level_rewards"].clone(), roll["response_mask"], roll["uid"])
    got, _ = st.compute_grpo_outcome_advantage(roll["token_level_rewards"].to(dev), roll["response_mask"].to(dev), roll["uid"])
    adv_close(got, want)
    # float and bool masks, n = 16
    roll = O.synth_rollout(512, 77, 1000, 16, seed=10, ragged=True)
    want, _ = O.compute_grpo_outcome_advantage(roll["token_level_rewards"].clone(), roll["response_mask"], roll["uid"])
    for m in (roll["response_mask"].float(), roll["response_mask"].bool()):
        got, _ = st.compute_grpo_outcome_advantage(roll["token_level_rewards"].to(dev), m.to(dev), roll["uid"])
        adv_close(got, want)
    with pytest.raises(AssertionError):
        st.compute_grpo_outcome_advantage(torch.ones(3, 2, device=dev), torch.ones(3, 2, device=dev),
                                          np.array(["a", "a", "b"], dtype=object))
    # the sharded entry (scores -> all-gather -> statistics -> local broadcast) with a single rank is the same function
    got, ret = st.core_algos.compute_grpo_outcome_advantage_sharded(roll["token_level_rewards"].to(dev),
                                                                    roll["response_mask"].to(dev), roll["uid"], 0)
    assert got is ret
    adv_close(got, want)


# ================================================================================================ fused lm_head
HEAD_CASES = [
    # rows, H, V, sigma_w, temperature
    (300, 256, 2176, 0.02, 1.0),       # ragged rows, vocab tail tile of 128
    (1000, 512, 151936, 0.16, 0.7),    # Qwen2.5 vocab, peaked logits (std ~ 5, max ~ 25)
    (515, 128, 152064, 0.3, 1.3),      # HF-padded 7B vocab, very peaked
    (129, 64, 1000, 0.05, 1.0),        # vocab not a multiple of 32 (masked tail columns)
]


@pytest.mark.parametrize("rows,h,v,sigma,temp", HEAD_CASES)
def test_fused_log_probs_forward(st, dev, rows, h, v, sigma, temp):
    hid, w = O.synth_head(rows, h, v, seed=rows, sigma_w=sigma)
    lab = torch.randint(0, v, (rows,), generator=torch.Generator().manual_seed(rows))
    want_lp, want_ent = O.lm_head_log_probs(hid, w, lab, temp, want_entropy=True)
    lp, ent = st.fused_lm_head_log_probs(hid.to(dev), w.to(dev), lab.to(dev), temp, want_entropy=True)
    assert float((lp.cpu() - want_lp).abs().max()) < TOL_LOGP
    assert float((ent.cpu() - want_ent).abs().max()) < TOL_LOGP
    assert bool((lp <= 0).all()) and bool((ent >= -1e-4).all()) and bool((ent <= math.log(v) + 1e-3).all())
    lp2, none = st.fused_lm_head_log_probs(hid.to(dev), w.to(dev), lab.to(dev), temp)
    assert none is None and torch.equal(lp2, lp)


@pytest.mark.parametrize("rows,h,v,sigma,temp", HEAD_CASES[:3])
def test_fused_log_probs_backward(st, dev, rows, h, v, sigma, temp):
    hid, w = O.synth_head(rows, h, v, seed=rows + 1, sigma_w=sigma)
    g = torch.Generator().manual_seed(rows)
    lab = torch.randint(0, v, (rows,), generator=g)
    gl = torch.randn(rows, generator=g) / rows
    gl[::5] = 0.0  # masked tokens
    ge = torch.randn(rows, generator=g) / rows
    hf, wf = hid.float().requires_grad_(True), w.float().requires_grad_(True)
    lp_ref, ent_ref = O.lm_head_log_probs(hf, wf, lab, temp, want_entropy=True)
    ((lp_ref * gl).sum() + (ent_ref * ge).sum()).backward()
    hd, wd = hid.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True)
    lp, ent = st.fused_lm_head_log_probs(hd, wd, lab.to(dev), temp, want_entropy=True)
    ((lp * gl.to(dev)).sum() + (ent * ge.to(dev)).sum()).backward()
    assert hd.grad.dtype == torch.bfloat16 and wd.grad.dtype == torch.bfloat16
    assert rel(hd.grad, hf.grad) < TOL_REL
    assert rel(wd.grad, wf.grad) < TOL_REL
    assert float(hd.grad[::5].abs().max()) == 0.0 or rel(hd.grad[::5], hf.grad[::5]) < TOL_REL


def test_fused_log_probs_multi_chunk_consistency(st, dev):
    """More rows than one internal chunk (grpo_chunk_capacity_rows() = 18 944 by default): results must not depend on
    where the chunk boundary falls."""
    from spatialthinker_b200 import _lib

    cap = int(_lib.load().grpo_chunk_capacity_rows())
    rows, h, v = cap + 700, 128, 4096
    cut = cap - 500  # below the chunk boundary: the tail spans both chunks, with a different row-to-tile assignment
    hid, w = O.synth_head(rows, h, v, seed=5, sigma_w=0.1)
    lab = torch.randint(0, v, (rows,), generator=torch.Generator().manual_seed(5))
    hd, wd, ld = hid.to(dev), w.to(dev), lab.to(dev)
    whole, _ = st.fused_lm_head_log_probs(hd, wd, ld)
    tail, _ = st.fused_lm_head_log_probs(hd[cut:], wd, ld[cut:])
    assert torch.equal(whole[cut:], tail)  # bit-exact: a row's K loop does not depend on the tile it sits in
    want, _ = O.lm_head_log_probs(hid[cut:], w, lab[cut:])
    assert float((tail.cpu() - want).abs().max()) < TOL_LOGP
    # the same through a smaller chunk (three chunks, the boundary elsewhere): the option changes the schedule, not the values
    _lib.check(_lib.load().grpo_set_option(b"chunk_rows", 8192), "set_option")
    try:
        small, _ = st.fused_lm_head_log_probs(hd, wd, ld)
    finally:
        _lib.check(_lib.load().grpo_set_option(b"chunk_rows", 0), "set_option")
    assert float((small - whole).abs().max()) < 1e-5
    want_all, _ = O.lm_head_log_probs(hid, w, lab)
    assert float((small.cpu() - want_all).abs().max()) < TOL_LOGP


# ================================================================================================ fused GRPO loss
def _loss_inputs(bsz, tl, h, v, n, sigma, seed, ragged=True):
    hid, w = O.synth_head(bsz * tl, h, v, seed=seed, sigma_w=sigma)
    hid = hid.view(bsz, tl, h)
    roll = O.synth_rollout(bsz, tl, v, n, seed=seed, ragged=ragged)
    logp, _ = O.lm_head_log_probs(hid, w, roll["responses"])
    adv, _ = O.compute_grpo_outcome_advantage(roll["token_level_rewards"].clone(), roll["response_mask"], roll["uid"])
    return {"hidden": hid, "weight": w, "labels": roll["responses"], "mask": roll["response_mask"], "adv": adv,
            "old": O.perturbed_log_probs(logp, seed=seed + 1, outlier_frac=0.03),
            "ref": O.perturbed_log_probs(logp, seed=seed + 2, outlier_frac=0.03), "roll": roll}


def _check_fused_loss(st, dev, x, *, temperature=1.0, kl_penalty="low_var_kl", kl_coef=1e-2, grad_accum=4.0, use_ref=True,
                      entropy_coeff=0.0):
    want = O.fused_loss_reference(x["hidden"], x["weight"], x["labels"], x["old"], x["adv"], x["mask"],
                                  x["ref"] if use_ref else None, temperature=temperature, kl_penalty=kl_penalty,
                                  kl_coef=kl_coef, grad_accum=grad_accum, entropy_coef=entropy_coeff, want_entropy=True)
    hd = x["hidden"].to(dev).requires_grad_(True)
    wd = x["weight"].to(dev).requires_grad_(True)
    loss, met = st.fused_grpo_loss(hd, wd, x["labels"].to(dev), x["old"].to(dev), x["adv"].to(dev),
                                   x["ref"].to(dev) if use_ref else None, x["mask"].to(dev), temperature=temperature,
                                   clip_ratio_low=CLIP[0], clip_ratio_high=CLIP[1], clip_ratio_dual=CLIP[2],
                                   kl_penalty=kl_penalty, kl_coef=kl_coef, grad_accum=grad_accum,
                                   entropy_coeff=entropy_coeff, want_entropy=True)
    loss.backward()
    wm = want["metrics"]
    valid = x["mask"].bool()
    assert float((met["log_probs"].cpu() - want["log_probs"])[valid].abs().max()) < TOL_LOGP
    assert float((met["entropy"].cpu() - want["entropy"])[valid].abs().max()) < TOL_LOGP
    assert abs(float(loss) - float(want["loss"])) <= TOL_REL * abs(float(want["loss"])) + 1e-7
    for key in ("actor/pg_loss", "actor/entropy_loss", "actor/ppo_kl") + (("actor/kl_loss",) if use_ref else ()):
        assert abs(float(met[key]) - float(wm[key])) <= TOL_REL * abs(float(wm[key])) + 1e-5, key
    for key in ("actor/pg_clipfrac_higher", "actor/pg_clipfrac_lower"):
        assert abs(float(met[key]) - float(wm[key])) <= 2e-3, key  # a token on a clip boundary may flip
    assert rel(hd.grad, want["dhidden"]) < TOL_REL
    assert rel(wd.grad, want["dweight"]) < TOL_REL
    pad = ~valid
    if pad.any():
        assert float(hd.grad.cpu()[pad].abs().max()) == 0.0  # masked rows get exactly zero gradient
    return want, met


def test_fused_grpo_loss_golden(st, dev, golden):
    g = golden("end_to_end")
    for tag in ("flat", "peaked"):
        hid = t(g[f"{tag}_hidden"]).to(torch.bfloat16)
        w = t(g[f"{tag}_weight"]).to(torch.bfloat16)
        uid = np.array([str(u) for u in g[f"{tag}_uid"]], dtype=object)
        adv, _ = st.compute_grpo_outcome_advantage(t(g[f"{tag}_rewards"]).to(dev), t(g[f"{tag}_mask"]).to(dev), uid)
        adv_close(adv, t(g[f"{tag}_adv"]))
        hd, wd = hid.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True)
        loss, met = st.fused_grpo_loss(hd, wd, t(g[f"{tag}_labels"]).to(dev), t(g[f"{tag}_old"]).to(dev), adv,
                                       t(g[f"{tag}_ref"]).to(dev), t(g[f"{tag}_mask"]).to(dev),
                                       temperature=float(g[f"{tag}_temp"][0]), kl_penalty="low_var_kl", kl_coef=0.01,
                                       grad_accum=2.0, want_entropy=True)
        loss.backward()
        sc = g[f"{tag}_scalars"]  # total, cf_hi, cf_lo, entropy_loss, ppo_kl, kl, pg
        assert abs(float(loss) - sc[0] / 2.0) <= TOL_REL * abs(sc[0] / 2.0)
        assert abs(float(met["actor/pg_loss"]) - sc[0]) <= TOL_REL * abs(sc[0])
        assert abs(float(met["actor/pg_clipfrac_higher"]) - sc[1]) < 2e-2
        assert abs(float(met["actor/entropy_loss"]) - sc[3]) <= TOL_REL * abs(sc[3])
        assert abs(float(met["actor/kl_loss"]) - sc[5]) <= TOL_REL * abs(sc[5]) + 1e-6
        valid = t(g[f"{tag}_mask"]).bool()
        assert float((met["log_probs"].cpu() - t(g[f"{tag}_logp"]))[valid].abs().max()) < TOL_LOGP
        assert float((met["entropy"].cpu() - t(g[f"{tag}_entropy"]))[valid].abs().max()) < TOL_LOGP
        assert rel(hd.grad, t(g[f"{tag}_dhidden"])) < TOL_REL
        assert rel(wd.grad, t(g[f"{tag}_dweight"])) < TOL_REL


@pytest.mark.parametrize("sigma,temp", [(0.02, 1.0), (0.2, 0.8)])
def test_fused_grpo_loss_vs_oracle(st, dev, sigma, temp):
    x = _loss_inputs(8, 96, 256, 8192, 4, sigma, seed=31)
    _check_fused_loss(st, dev, x, temperature=temp)


@pytest.mark.parametrize("mode", ["kl", "abs", "mse", "chi2"])
def test_fused_grpo_loss_kl_modes(st, dev, mode):
    x = _loss_inputs(4, 40, 128, 2048, 4, 0.1, seed=41)
    _check_fused_loss(st, dev, x, kl_penalty=mode, kl_coef=0.05)


def test_fused_grpo_loss_no_ref_and_entropy_bonus(st, dev):
    x = _loss_inputs(4, 64, 128, 4096, 4, 0.1, seed=51)
    _check_fused_loss(st, dev, x, use_ref=False)
    _check_fused_loss(st, dev, x, entropy_coeff=0.01)


def test_fused_grpo_loss_edge_cases(st, dev):
    x = _loss_inputs(4, 32, 128, 2048, 4, 0.1, seed=61)
    hd, wd = x["hidden"].to(dev).requires_grad_(True), x["weight"].to(dev).requires_grad_(True)
    zero = torch.zeros_like(x["mask"])
    loss, met = st.fused_grpo_loss(hd, wd, x["labels"].to(dev), x["old"].to(dev), x["adv"].to(dev), None, zero.to(dev))
    loss.backward()
    assert float(loss) == 0.0 and float(hd.grad.abs().max()) == 0.0 and float(wd.grad.abs().max()) == 0.0  # KAT-D
    # on-policy: old == current log-probs => ratio 1, no clipping, ppo_kl 0
    lp, _ = st.fused_lm_head_log_probs(x["hidden"].to(dev), x["weight"].to(dev), x["labels"].to(dev))
    loss, met = st.fused_grpo_loss(x["hidden"].to(dev), x["weight"].to(dev), x["labels"].to(dev), lp, x["adv"].to(dev),
                                   None, x["mask"].to(dev))
    assert float(met["actor/pg_clipfrac_higher"]) == 0 and float(met["actor/pg_clipfrac_lower"]) == 0
    assert abs(float(met["actor/ppo_kl"])) < 1e-7
    with pytest.raises(ValueError):
        st.fused_grpo_loss(x["hidden"].float().to(dev), x["weight"].to(dev), x["labels"].to(dev), lp, x["adv"].to(dev),
                           None, x["mask"].to(dev))
    with pytest.raises(NotImplementedError):
        st.fused_grpo_loss(x["hidden"].to(dev), x["weight"].to(dev), x["labels"].to(dev), lp, x["adv"].to(dev),
                           x["ref"].to(dev), x["mask"].to(dev), kl_penalty="full", kl_coef=0.1)


def test_improbable_labels_and_garbage_padding(st, dev):
    """The softmax is referenced to the label's own logit. Labels of very low probability (log p ~ -45) must still be
    exact, and padded rows whose label is absurdly improbable must stay finite and contribute exactly nothing."""
    bsz, tl, h, v = 4, 48, 128, 4096
    g = torch.Generator().manual_seed(7)
    hid = torch.randn(bsz, tl, h, generator=g).to(torch.bfloat16)
    w = (0.45 * torch.randn(v, h, generator=g)).to(torch.bfloat16)  # logit std ~ 5, spread ~ +-20
    z = hid.float() @ w.float().t()
    labels = z.argmin(-1)  # the least likely token of every row
    labels[:, ::3] = z.argmax(-1)[:, ::3]
    want_lp, want_ent = O.lm_head_log_probs(hid, w, labels, 0.7, want_entropy=True)  # T = 0.7 widens the spread
    assert -66 < float(want_lp.min()) < -40  # exact down to log p = -69 (kClampLog2), far below anything sampled
    lp, ent = st.fused_lm_head_log_probs(hid.to(dev), w.to(dev), labels.to(dev), 0.7, want_entropy=True)
    assert float((lp.cpu() - want_lp).abs().max()) < TOL_LOGP  # north_star: 2e-3 absolute, also at log p = -65
    assert float((ent.cpu() - want_ent).abs().max()) < TOL_LOGP
    # garbage padding: sequences 1 and 3 are fully masked and point at a token that is > 100 nats below the row maximum
    # at position 0 (the clamp binds there); the valid sequences keep ordinary labels
    w2 = w.clone()
    w2[7] = (-8.0 * hid[1, 0].float() / hid[1, 0].float().norm() * 1.0).to(torch.bfloat16)
    w2[7] += (-8.0 * hid[3, 0].float() / hid[3, 0].float().norm()).to(torch.bfloat16)
    z2 = hid.float() @ w2.float().t()
    assert float(z2[1, 0].max() - z2[1, 0, 7]) > 80
    mask = torch.ones(bsz, tl, dtype=torch.int64)
    mask[1] = 0
    mask[3] = 0
    labels2 = z2.argmax(-1)
    labels2[:, 1::2] = torch.randint(0, v, (bsz, tl // 2), generator=g)
    labels2[1] = 7
    labels2[3] = 7
    x = {"hidden": hid, "weight": w2, "labels": labels2, "mask": mask, "adv": torch.randn(bsz, 1, generator=g).expand(bsz, tl).contiguous()}
    lp_ref, _ = O.lm_head_log_probs(hid, w2, labels2, 1.0)
    assert float(lp_ref[mask.bool()].min()) > -60
    x["old"] = O.perturbed_log_probs(lp_ref, seed=3)
    x["ref"] = O.perturbed_log_probs(lp_ref, seed=4)
    want, met = _check_fused_loss(st, dev, x, temperature=1.0)
    assert bool(torch.isfinite(met["log_probs"]).all())


@pytest.mark.parametrize("n", [1, 37, 2048, 2049, 300001])
@pytest.mark.parametrize("mdtype", [torch.int64, torch.float32, torch.bool])
def test_compaction_kernels(st, dev, n, mdtype):
    """Stable compaction index / row gather / row scatter (csrc/compact_kernels.cuh) against torch on the same mask;
    all-masked and all-valid inputs included. Bit-exact."""
    from spatialthinker_b200 import fused

    g = torch.Generator().manual_seed(n)
    for density in (0.0, 0.37, 1.0):
        mask = (torch.rand(n, generator=g) < density).to(mdtype).to(dev)
        gather_idx, inverse, count = fused.compact_index(mask)
        want = torch.nonzero(mask.reshape(-1) != 0).squeeze(1)
        m = int(count.item())
        assert m == want.numel()
        assert torch.equal(gather_idx[:m].long(), want)
        inv_want = torch.full((n,), -1, dtype=torch.int32, device=dev)
        inv_want[want] = torch.arange(m, dtype=torch.int32, device=dev)
        assert torch.equal(inverse, inv_want)
        for shape, dt in (((n, 24), torch.bfloat16), ((n, 130), torch.bfloat16), ((n,), torch.float32), ((n,), torch.int64)):
            src = torch.randn(shape, generator=g).mul(100).to(dt).to(dev)
            got = fused.gather_rows(src, gather_idx, m)
            assert torch.equal(got, src[want])
            back = fused.scatter_rows(got, inverse)
            ref = torch.zeros_like(src)
            ref[want] = src[want]
            assert torch.equal(back, ref)


def test_padding_compaction_is_equivalent(st, dev):
    """valid_rows (host-known count of unmasked tokens) drops padded rows before the GEMMs; nothing else may change."""
    x = _loss_inputs(8, 96, 128, 4096, 4, 0.1, seed=131, ragged=True)
    d = {k: x[k].to(dev) for k in ("hidden", "weight", "labels", "old", "adv", "ref", "mask")}
    n_valid = int(x["mask"].sum())
    assert 0 < n_valid < x["mask"].numel()
    kw = dict(temperature=0.9, kl_penalty="low_var_kl", kl_coef=0.02, grad_accum=2.0, want_entropy=True)
    full = st.grpo_micro_batch_step(d["hidden"], d["weight"], d["labels"], d["old"], d["adv"], d["ref"], d["mask"], **kw)
    comp = st.grpo_micro_batch_step(d["hidden"], d["weight"], d["labels"], d["old"], d["adv"], d["ref"], d["mask"],
                                    valid_rows=n_valid, **kw)
    valid = x["mask"].bool().to(dev)
    np.testing.assert_allclose(comp["metrics"].cpu().numpy(), full["metrics"].cpu().numpy(), rtol=2e-5, atol=1e-7)
    assert float((comp["log_probs"] - full["log_probs"])[valid].abs().max()) < 1e-5
    assert float(comp["log_probs"][~valid].abs().max()) == 0.0
    assert rel(comp["dhidden"], full["dhidden"]) < 2e-3 and float(comp["dhidden"][~valid].abs().max()) == 0.0
    assert rel(comp["dweight"], full["dweight"]) < 2e-3
    # and against the oracle
    want = O.fused_loss_reference(x["hidden"], x["weight"], x["labels"], x["old"], x["adv"], x["mask"], x["ref"],
                                  temperature=0.9, kl_penalty="low_var_kl", kl_coef=0.02, grad_accum=2.0)
    assert rel(comp["dhidden"], want["dhidden"]) < TOL_REL and rel(comp["dweight"], want["dweight"]) < TOL_REL
    # the actor loop uses it by default: same metrics with and without
    cfg = st.ActorConfig(global_batch_size_per_device=8, micro_batch_size_per_device_for_update=4, use_kl_loss=True,
                         kl_penalty="low_var_kl", kl_coef=0.02)
    batch = {"hidden_states": d["hidden"], "responses": d["labels"], "response_mask": d["mask"], "old_log_probs": d["old"],
             "advantages": d["adv"], "ref_log_probs": d["ref"]}
    m1 = st.DataParallelPPOActor(cfg, d["weight"], compact_padding=True).update_policy(st.TensorBatch(batch, meta_info={"temperature": 0.9}))
    m0 = st.DataParallelPPOActor(cfg, d["weight"], compact_padding=False).update_policy(st.TensorBatch(batch, meta_info={"temperature": 0.9}))
    for key in ("actor/pg_loss", "actor/ppo_kl", "actor/entropy_loss", "actor/grad_norm"):
        np.testing.assert_allclose(m1[key], m0[key], rtol=2e-3, atol=1e-6)


def test_config_c1_full_head(st, dev):
    """BASELINE.json configs[0]: Qwen2.5-VL-3B head (H 2048, V 151936), 8 rollouts x 512 response tokens, against the
    fp32 CPU oracle - log-probs, entropy, advantages, loss, dHidden, dW."""
    x = _loss_inputs(8, 512, 2048, 151936, 8, 0.02, seed=71)
    got_adv, _ = st.compute_grpo_outcome_advantage(x["roll"]["token_level_rewards"].to(dev), x["mask"].to(dev), x["roll"]["uid"])
    adv_close(got_adv, x["adv"])
    _check_fused_loss(st, dev, x, grad_accum=1.0)


# ================================================================================================ actor loop
def test_update_policy_matches_reference_loop(st, dev):
    bsz, tl, h, v = 16, 48, 128, 4096
    x = _loss_inputs(bsz, tl, h, v, 4, 0.1, seed=81)
    batch = {"responses": x["labels"], "response_mask": x["mask"], "advantages": x["adv"], "old_log_probs": x["old"],
             "ref_log_probs": x["ref"]}
    want = O.update_policy_reference(x["hidden"], x["weight"], batch, global_batch_size_per_device=8,
                                     micro_batch_size_per_device_for_update=2, temperature=0.9, kl_penalty="low_var_kl",
                                     kl_coef=0.01)
    cfg = st.ActorConfig(global_batch_size_per_device=8, micro_batch_size_per_device_for_update=2,
                         micro_batch_size_per_device_for_experience=4, use_kl_loss=True, kl_penalty="low_var_kl", kl_coef=0.01)
    actor = st.DataParallelPPOActor(cfg, x["weight"].to(dev))
    tensors = {k: val.to(dev) for k, val in batch.items()}
    tensors["hidden_states"] = x["hidden"].to(dev)
    # the mask can also arrive as the tail of attention_mask, as in the reference (dp_actor.py:247)
    data = st.TensorBatch(tensors, meta_info={"temperature": 0.9})
    lp = actor.compute_log_prob(data)
    want_lp, _ = O.lm_head_log_probs(x["hidden"], x["weight"], x["labels"], 0.9)
    valid = x["mask"].bool()
    # padded slots are dropped before the GEMM (compact_padding, the default) and read 0; without it every slot is computed
    assert lp.shape == (bsz, tl) and float((lp.cpu() - want_lp)[valid].abs().max()) < TOL_LOGP
    assert float(lp.cpu()[~valid].abs().max()) == 0.0
    full = st.DataParallelPPOActor(cfg, x["weight"].to(dev), compact_padding=False).compute_log_prob(data)
    assert float((full.cpu() - want_lp).abs().max()) < TOL_LOGP
    dws = []
    orig_step = actor._optimizer_step

    def spy():
        dws.append(actor.dweight.clone())
        return orig_step()

    actor._optimizer_step = spy
    met = actor.update_policy(data)
    assert len(dws) == 2 and len(met["actor/pg_loss"]) == 8 and len(met["actor/grad_norm"]) == 2
    assert isinstance(met["actor/kl_loss"], float) and met["actor/kl_coef"] == 0.01
    for key in ("actor/pg_loss", "actor/entropy_loss", "actor/ppo_kl"):
        np.testing.assert_allclose(met[key], want["metrics"][key], rtol=TOL_REL, atol=1e-5)
    for i in range(2):
        assert rel(dws[i], want["steps"][i]["dweight"]) < TOL_REL
        gn = float(want["steps"][i]["dweight"].norm())
        assert abs(met["actor/grad_norm"][i] - gn) <= TOL_REL * gn
    got_dh = torch.cat([d for d in actor.last_dhidden], dim=0)
    want_dh = want["steps"][0]["dhidden"] + want["steps"][1]["dhidden"]
    assert rel(got_dh, want_dh) < TOL_REL
    # attention_mask form of the mask
    tensors2 = dict(tensors)
    del tensors2["response_mask"]
    tensors2["attention_mask"] = torch.cat([torch.ones(bsz, 5, dtype=torch.int64, device=dev), x["mask"].to(dev)], dim=1)
    met2 = actor.update_policy(st.TensorBatch(tensors2, meta_info={"temperature": 0.9}))
    np.testing.assert_allclose(met2["actor/pg_loss"], met["actor/pg_loss"], rtol=1e-6)


def test_deferred_dw_matches_per_micro_batch(st, dev):
    """fused.DeferredDW: several small micro-batches share one chunk workspace (each in its own 512-row-aligned slot) and
    ONE dW GEMM at flush time. Log-probs and metrics are the per-micro-batch ones bit for bit (same kernels on the same
    rows), dHidden and dW agree up to fp32 summation order, and dW matches the oracle's accumulated gradient. Slot
    sizes are ragged so that the gaps up to the next 512-row boundary are exercised; the last micro-batch no longer
    fits and forces a flush in the middle."""
    from spatialthinker_b200 import _lib
    from spatialthinker_b200.fused import DeferredDW

    lib = _lib.load()
    h, v = 256, 4096 + 520
    _lib.check(lib.grpo_set_option(b"chunk_rows", 4096), "set_option")  # a small workspace: 4 slots, then it is full
    try:
        assert lib.grpo_chunk_capacity_rows() == 4096
        sizes = [(7, 100), (4, 128), (9, 140), (1, 90), (6, 200)]  # (sequences, tokens) -> 700, 512, 1260, 90, 1200 rows
        _, w = O.synth_head(8, h, v, seed=200, sigma_w=0.1)
        cases = []
        for i, (b, tl) in enumerate(sizes):  # every micro-batch against the SAME weight
            hid, _ = O.synth_head(b * tl, h, v, seed=201 + i, sigma_w=0.1)
            hid = hid.view(b, tl, h)
            roll = O.synth_rollout(b, tl, v, b, seed=201 + i, ragged=(i % 2 == 0))
            logp, _ = O.lm_head_log_probs(hid, w, roll["responses"], 0.9)
            g = torch.Generator().manual_seed(300 + i)
            adv = torch.randn(b, 1, generator=g).expand(b, tl) * roll["response_mask"]
            cases.append({"hidden": hid, "labels": roll["responses"], "mask": roll["response_mask"], "adv": adv.contiguous(),
                          "old": O.perturbed_log_probs(logp, seed=400 + i, outlier_frac=0.03),
                          "ref": O.perturbed_log_probs(logp, seed=500 + i, outlier_frac=0.03)})
        wd = w.to(dev)
        kw = dict(temperature=0.9, kl_penalty="low_var_kl", kl_coef=0.01, grad_accum=float(len(cases)),
                  clip_ratio_low=CLIP[0], clip_ratio_high=CLIP[1], clip_ratio_dual=CLIP[2], want_entropy=True)

        def run(deferred):
            dw = torch.zeros(v, h, dtype=torch.float32, device=dev)
            session = DeferredDW(wd, dw) if deferred else None
            outs, row0s = [], []
            for x in cases:
                res = st.grpo_micro_batch_step(x["hidden"].to(dev), wd, x["labels"].to(dev), x["old"].to(dev),
                                               x["adv"].to(dev), x["ref"].to(dev), x["mask"].to(dev), dweight_accum=dw,
                                               defer=session, **kw)
                outs.append((res["log_probs"].clone(), res["entropy"].clone(), res["metrics"].clone(), res["dhidden"].clone()))
                if session is not None:
                    row0s.append(session.total_rows - x["labels"].numel())
            if session is not None:
                assert session.pending == 1  # the fifth micro-batch did not fit: flushed, then slot 0 again
                session.flush()
                assert session.pending == 0 and session.total_rows == 0
            torch.cuda.synchronize()
            return outs, dw, row0s

        base, dw_base, _ = run(False)
        got, dw_got, row0s = run(True)
        assert row0s == [0, 1024, 1536, 3072, 0]
        for (lp0, e0, m0, dh0), (lp1, e1, m1, dh1) in zip(base, got):
            assert torch.equal(lp0, lp1) and torch.equal(e0, e1) and torch.equal(m0, m1)
            assert rel(dh1, dh0) < 2e-3
        assert rel(dw_got, dw_base) < 1e-4
        want_dw = torch.zeros(v, h)
        for x in cases:
            want = O.fused_loss_reference(x["hidden"], w, x["labels"], x["old"], x["adv"], x["mask"], x["ref"],
                                          temperature=0.9, kl_penalty="low_var_kl", kl_coef=0.01,
                                          grad_accum=float(len(cases)))
            want_dw += want["dweight"]
        assert rel(dw_got, want_dw) < TOL_REL
    finally:
        lib.grpo_set_option(b"chunk_rows", 0)


def test_update_policy_deferred_dw(st, dev):
    """DataParallelPPOActor(defer_dw=True): same metrics, same dW per optimizer step, same dHidden as the default loop."""
    bsz, tl, h, v = 16, 48, 128, 4096
    x = _loss_inputs(bsz, tl, h, v, 4, 0.1, seed=81, ragged=True)
    cfg = st.ActorConfig(global_batch_size_per_device=8, micro_batch_size_per_device_for_update=2,
                         micro_batch_size_per_device_for_experience=4, use_kl_loss=True, kl_penalty="low_var_kl", kl_coef=0.01)
    tensors = {"responses": x["labels"].to(dev), "response_mask": x["mask"].to(dev), "advantages": x["adv"].to(dev),
               "old_log_probs": x["old"].to(dev), "ref_log_probs": x["ref"].to(dev), "hidden_states": x["hidden"].to(dev)}
    data = st.TensorBatch(tensors, meta_info={"temperature": 0.9})
    res = []
    for defer in (False, True):
        actor = st.DataParallelPPOActor(cfg, x["weight"].to(dev), defer_dw=defer)
        dws = []
        orig = actor._optimizer_step

        def spy(actor=actor, dws=dws, orig=orig):
            dws.append(actor.dweight.clone())
            return orig()

        actor._optimizer_step = spy
        met = actor.update_policy(data)
        res.append((met, dws, torch.cat(list(actor.last_dhidden), dim=0)))
    (m0, dw0, dh0), (m1, dw1, dh1) = res
    assert len(dw0) == len(dw1) == 2
    for key in ("actor/pg_loss", "actor/entropy_loss", "actor/ppo_kl", "actor/pg_clipfrac_higher"):
        assert m0[key] == m1[key], key
    for a, b in zip(dw0, dw1):
        assert rel(b, a) < 1e-4
    np.testing.assert_allclose(m1["actor/grad_norm"], m0["actor/grad_norm"], rtol=1e-4)
    assert rel(dh1, dh0) < 2e-3


def test_update_policy_with_hidden_fn_and_optimizer(st, dev):
    """hidden_fn keeps an autograd graph: dHidden must reach the parameters behind it, and the optimizer must move W."""
    bsz, tl, h, v = 8, 16, 64, 1024
    x = _loss_inputs(bsz, tl, h, v, 4, 0.1, seed=91)
    body = torch.nn.Linear(h, h, bias=False).to(dev).to(torch.bfloat16)
    weight = torch.nn.Parameter(x["weight"].to(dev).clone())
    opt = torch.optim.SGD([weight], lr=1.0)
    cfg = st.ActorConfig(global_batch_size_per_device=8, micro_batch_size_per_device_for_update=4)
    actor = st.DataParallelPPOActor(cfg, weight, actor_optimizer=opt, hidden_fn=lambda mb: body(mb["inputs"]))
    data = st.TensorBatch({"inputs": x["hidden"].to(dev), "responses": x["labels"].to(dev), "response_mask": x["mask"].to(dev),
                           "old_log_probs": x["old"].to(dev), "advantages": x["adv"].to(dev)}, meta_info={"temperature": 1.0})
    before = weight.detach().clone()
    met = actor.update_policy(data)
    assert body.weight.grad is not None and float(body.weight.grad.abs().sum()) > 0
    assert not torch.equal(before, weight.detach())
    assert math.isfinite(met["actor/grad_norm"][0])


# ================================================================================================ full-size properties
def test_full_size_properties_7b_head(st, dev):
    """BASELINE size (7B head, H 3584, V 151936) - too big for the CPU oracle in a test, so size-independent properties:
    additivity of dW over row blocks, exact zeros for masked rows, log p <= 0, 0 <= entropy <= ln V, on-policy ratio."""
    rows, h, v = 2 * 9472 + 1000, 3584, 151936  # three internal chunks, the last one ragged
    g = torch.Generator().manual_seed(123)
    hid = torch.randn(rows, h, generator=g).to(torch.bfloat16).to(dev)
    w = (0.02 * torch.randn(v, h, generator=g)).to(torch.bfloat16).to(dev)
    lab = torch.randint(0, v, (rows,), generator=g).to(dev)
    mask = (torch.rand(rows, generator=g) > 0.3).long().to(dev)
    adv = torch.randn(rows, generator=g).to(dev)
    lp, ent = st.fused_lm_head_log_probs(hid, w, lab, 1.0, want_entropy=True)
    assert bool((lp <= 0).all()) and bool((ent >= 0).all()) and bool((ent <= math.log(v) + 1e-3).all())
    # spot-check 64 rows against torch on the device in fp32 (logits for 64 rows only)
    idx = torch.randperm(rows, generator=g)[:64].to(dev)
    z = hid[idx].float() @ w.float().t()
    want = z.gather(1, lab[idx, None]).squeeze(1) - torch.logsumexp(z, -1)
    assert float((lp[idx] - want).abs().max()) < TOL_LOGP
    old = (lp + 0.05 * torch.randn(rows, generator=g).to(dev))
    kw = dict(temperature=1.0, kl_penalty=None, grad_accum=2.0)
    full = st.grpo_micro_batch_step(hid, w, lab, old, adv, None, mask, **kw)
    assert float(full["dhidden"][mask == 0].abs().max()) == 0.0
    # additivity: with the normaliser held fixed (same sum(mask) via grad_accum rescale), dW(all) = dW(A) + dW(B)
    cut = 9472 + 333
    m_all = float(mask.sum())
    parts = torch.zeros_like(full["dweight"])
    for sl in (slice(0, cut), slice(cut, rows)):
        ga = 2.0 * m_all / float(mask[sl].sum())  # loss_part / ga == contribution of the part to loss_all / 2
        st.grpo_micro_batch_step(hid[sl], w, lab[sl], old[sl], adv[sl], None, mask[sl], dweight_accum=parts,
                                 temperature=1.0, kl_penalty=None, grad_accum=ga)
    assert rel(parts, full["dweight"]) < 5e-3
    # on-policy
    on = st.grpo_micro_batch_step(hid, w, lab, lp, adv, None, mask, need_grads=False, **kw)
    m = on["metrics"].cpu()
    assert float(m[1]) == 0 and float(m[2]) == 0 and abs(float(m[3])) < 1e-7
