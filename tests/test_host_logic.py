"""CPU-side checks: the C ABI library loads and exports everything the header declares, host-side index logic,
containers, and the guarantee that the product never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from spatialthinker_b200 import _lib, build

    build.build(verbose=False)
    return _lib.load()


def test_header_symbols_exported(lib):
    from spatialthinker_b200 import _lib

    header = open(os.path.join(ROOT, "include", "grpo_b200.h")).read()
    declared = set(re.findall(r"\b(grpo_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/grpo_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.grpo_abi_version() == 2


def test_workspace_queries_are_host_only(lib):
    small = lib.grpo_lmhead_fwd_workspace_bytes(128, 2048, 151936)
    big = lib.grpo_fused_loss_workspace_bytes(1 << 20, 3584, 151936)
    assert 0 < small < big
    # the stash never exceeds one chunk of rows, whatever the micro-batch size
    assert big == lib.grpo_fused_loss_workspace_bytes(1 << 22, 3584, 151936)
    assert big < 6.5e9  # one chunk of 18944 rows: 5.8 GB of exp-stash + 0.14 GB scaled hidden + partial sums


def test_sass_is_blackwell_native():
    from spatialthinker_b200 import _lib

    try:
        sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):  # tcgen05.mma, TMA load, tcgen05.ld
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", "")  # no legacy mma.sync path


def test_product_fails_loudly_without_cuda_tensors():
    import spatialthinker_b200 as st

    z = torch.zeros(2, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        st.log_probs_from_logits(z, torch.zeros(2, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        st.compute_kl(z, z, "kl")
    with pytest.raises(NotImplementedError, match="Unknown KL penalty"):
        st.compute_kl(z, z, "full")


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from spatialthinker_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GrpoLibraryError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_no_oracle_in_product():
    pkg = os.path.join(ROOT, "spatialthinker_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "grpo_oracle" not in text, f


def test_group_csr():
    from spatialthinker_b200.core_algos import group_csr

    uid = np.array(["b", "a", "b", "c", "a", "c", "a"], dtype=object)
    order, offsets = group_csr(uid)
    groups = [sorted(order[offsets[i]:offsets[i + 1]].tolist()) for i in range(len(offsets) - 1)]
    assert sorted(groups) == [[0, 2], [1, 4, 6], [3, 5]]
    assert order.dtype == np.int32 and offsets.dtype == np.int32 and offsets[-1] == 7
    with pytest.raises(AssertionError, match="rollout.n > 1"):
        group_csr(np.array(["a", "a", "b"], dtype=object))
    order, offsets = group_csr(torch.tensor([3, 3, 9, 9]))
    assert offsets.tolist() == [0, 2, 4]


def _reference_partitions():
    """The reference's partitioner, executed from its source text (its module imports tensordict, which is absent)."""
    path = "/root/reference/verl/utils/seqlen_balancing.py"
    if not os.path.exists(path):
        pytest.skip("reference checkout not present")
    src = open(path).read()
    src = src.replace("from tensordict import TensorDict", "TensorDict = object")
    ns = {}
    exec(compile(src, path, "exec"), ns)
    return ns["get_seqlen_balanced_partitions"]


@pytest.mark.parametrize("equal", [True, False])
def test_balanced_partitions_match_reference(equal):
    from spatialthinker_b200.sharding import balanced_partitions

    ref = _reference_partitions()
    rng = np.random.default_rng(0)
    for n, k in ((64, 8), (32, 4), (128, 8), (16, 2), (24, 8)):
        lens = rng.integers(1, 8192, size=n).tolist()
        assert balanced_partitions(lens, k, equal) == ref(lens, k, equal)
    lens = [5] * 16  # all ties
    assert balanced_partitions(lens, 4, equal) == ref(lens, 4, equal)


def test_balanced_partitions_properties():
    from spatialthinker_b200.sharding import balanced_partitions, rank_rows

    rng = np.random.default_rng(1)
    lens = rng.integers(1, 4096, size=4096).tolist()
    parts = balanced_partitions(lens, 8, True)
    assert sorted(i for p in parts for i in p) == list(range(4096))
    assert all(len(p) == 512 for p in parts)
    sums = [sum(lens[i] for i in p) for p in parts]
    assert max(sums) - min(sums) <= 0.001 * max(sums)
    assert rank_rows(lens, 8, 3) == parts[3]
    with pytest.raises(AssertionError):
        balanced_partitions([1, 2, 3], 2, True)


def test_tensor_batch_split_semantics():
    from spatialthinker_b200.protocol import TensorBatch

    tb = TensorBatch({"a": torch.arange(12).view(6, 2), "b": torch.arange(6)}, {"uid": np.arange(6)}, {"temperature": 0.7})
    parts = tb.split(2)
    assert len(parts) == 3 and parts[1].batch["b"].tolist() == [2, 3] and parts[2].non_tensor_batch["uid"].tolist() == [4, 5]
    assert parts[0].meta_info["temperature"] == 0.7
    assert list(tb.select(["b"]).batch) == ["b"]
    with pytest.raises(AssertionError):
        tb.chunk(4)
    with pytest.raises(ValueError):
        TensorBatch({"a": torch.zeros(3), "b": torch.zeros(4)})


def test_patch_verl_roundtrip(reference_modules):
    import spatialthinker_b200 as st

    VF, ca = reference_modules
    orig = ca.compute_policy_loss
    done = st.patch_verl()
    try:
        assert "compute_policy_loss" in done["verl.trainer.core_algos"]
        assert ca.compute_policy_loss is not orig and ca.compute_policy_loss.__wrapped__ is orig
        if not torch.cuda.is_available():  # a GPU-less driver keeps running the reference's own code
            out = ca.compute_kl(torch.zeros(3), torch.ones(3), "kl")
            assert out.tolist() == [-1.0, -1.0, -1.0]
    finally:
        st.unpatch_verl()
    assert ca.compute_policy_loss is orig


def test_header_is_plain_c_and_links(lib, tmp_path):
    """The boundary is a C ABI: include/grpo_b200.h must compile as C99 and a plain C program must link against the
    library and get answers from its host-only entry points (no CUDA device needed for these)."""
    from spatialthinker_b200 import _lib

    src = tmp_path / "abi.c"
    src.write_text(
        '#include "grpo_b200.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  size_t ws = grpo_fused_loss_workspace_bytes(37888, 3584, 151936);\n"
        "  int rc = grpo_compute_kl(0, 0, 4, GRPO_KL_LOW_VAR, 0, 0, 0); /* null pointers: argument error, not a crash */\n"
        '  printf("%d %zu %d %s\\n", grpo_abi_version(), ws, rc, grpo_last_error());\n'
        "  return 0;\n}\n")
    exe = tmp_path / "abi"
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                         "-o", str(exe), _lib.LIB_PATH, "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    version, ws, rc, msg = out.stdout.split(maxsplit=3)
    assert int(version) == 2 and int(ws) == lib.grpo_fused_loss_workspace_bytes(37888, 3584, 151936)
    assert int(rc) == -1 and "null" in msg


def _plan(lib, tiles, k_blocks, groups, mode):
    info = (ctypes.c_int32 * 4)()
    assert lib.grpo_debug_plan_units(tiles, k_blocks, groups, mode, None, 0, info) == 0
    n_groups, n_units, bound, slices = list(info)
    units = (ctypes.c_int32 * (3 * n_units))()
    assert lib.grpo_debug_plan_units(tiles, k_blocks, groups, mode, units, n_units, info) == 0
    return n_groups, np.array(units, dtype=np.int64).reshape(n_units, 3), bound, slices


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_work_unit_plan_covers_every_k_block_once(lib, mode):
    """The persistent GEMMs walk `units` (whole tiles + split-K slices of the last partial round): every (tile, K-block)
    must be covered exactly once, and no group may load more K-blocks than the progress-window bound handed to the
    producers (an underestimate would leave a producer waiting for arrivals that never come)."""
    rng = np.random.default_rng(5)
    cases = [(112, 2374, 74), (28, 2374, 74), (4158, 296, 74), (518, 2374, 74), (1, 32, 74), (75, 2, 74), (80, 1, 74),
             (149, 24, 148), (3, 7, 2), (73, 17, 74)]
    cases += [(int(rng.integers(1, 700)), int(rng.integers(1, 300)), int(rng.integers(1, 149))) for _ in range(60)]
    for tiles, kb, groups in cases:
        n_groups, units, bound, slices = _plan(lib, tiles, kb, groups, mode)
        assert 1 <= n_groups <= groups
        cover = np.zeros((tiles, kb), dtype=np.int32)
        load = np.zeros(n_groups, dtype=np.int64)
        for u, (t, k0, k1) in enumerate(units):
            assert 0 <= t < tiles and 0 <= k0 < k1 <= kb, (tiles, kb, groups, u, t, k0, k1)
            cover[t, k0:k1] += 1
            load[u % n_groups] += k1 - k0
        assert (cover == 1).all(), (tiles, kb, groups, mode)
        assert load.max() <= bound, (tiles, kb, groups, mode, load.max(), bound)
        if mode == 0:
            assert slices == 0 and len(units) == tiles
        if mode == 1 and slices:  # one short extra round
            assert len(units) - (tiles - tiles % n_groups) <= n_groups


def test_split_plan_shortens_the_reference_micro_batch(lib):
    """dHidden GEMM of the reference's shipped micro-batch (4 sequences x 1024 tokens, 7B head: 8 x 14 wide tiles on 74
    CTA pairs = 1.51 rounds): whole tiles take 2 rounds, the split plan stays within 10 % of the ideal 1.51."""
    tiles, kb, groups = 8 * 14, 2374, 74
    n_groups, units, _, slices = _plan(lib, tiles, kb, groups, 2)
    assert n_groups == groups and slices > 1
    load = np.zeros(groups, dtype=np.int64)
    for u, (_, k0, k1) in enumerate(units):
        load[u % groups] += k1 - k0
    assert load.max() < 1.10 * tiles * kb / groups
    assert load.max() < 0.85 * 2 * kb
    # slice-major tail: the groups running side by side in a tail round work on (at most two) common K ranges
    tail = units[tiles - tiles % groups:]
    first_round = tail[:groups]
    assert len({(k0, k1) for _, k0, k1 in first_round.tolist()}) <= 2
    # fewer tiles than groups: every group gets work
    n_groups, units, _, slices = _plan(lib, 28, kb, groups, 2)
    assert n_groups == groups and slices > 1 and len(units) > groups


def test_deferred_dw_slot_bookkeeping():
    """fused.DeferredDW.reserve: slots start on 512-row boundaries, a micro-batch that no longer fits flushes first, one
    larger than the workspace is refused (the caller then takes the ordinary path). Pure host arithmetic."""
    from spatialthinker_b200.fused import DeferredDW

    s = object.__new__(DeferredDW)  # no device: only the bookkeeping is exercised
    s.capacity, s.next_row0, s.total_rows, s.pending = 18944, 0, 0, 0
    s.workspace = object()  # stands in for the lazily allocated device buffer
    flushed = []

    def flush():
        flushed.append((s.total_rows, s.pending))
        s.next_row0 = s.total_rows = s.pending = 0

    s.flush = flush
    assert [s.reserve(n) for n in (4096, 4000, 100, 4096)] == [0, 4096, 8192, 8704]
    assert (s.total_rows, s.next_row0, s.pending) == (12800, 12800, 4) and not flushed
    assert s.reserve(6144) == 12800 and s.total_rows == 18944  # exactly full
    assert s.reserve(1) == 0 and flushed == [(18944, 5)]  # did not fit: flushed, first slot again
    assert s.reserve(18945) is None and s.reserve(0) is None  # never fits / empty: ordinary path
    assert (s.total_rows, s.next_row0, s.pending) == (1, 512, 1)


def test_rearrange_micro_batches_matches_reference():
    """sharding.rearrange_micro_batches = verl/utils/seqlen_balancing.py:222-255 on host-known lengths: the same number of
    micro-batches, and - when the reference checkout is present - the very same partitions."""
    import random

    from spatialthinker_b200.protocol import TensorBatch
    from spatialthinker_b200.sharding import rearrange_micro_batches

    rng = random.Random(4)
    cases = [([rng.randint(1, 4096) for _ in range(n)], cap) for n, cap in ((16, 8192), (128, 37888), (7, 4096), (64, 5000))]
    cases.append(([1024] * 128, 37888))
    ref = None
    ref_root = os.environ.get("GRPO_REFERENCE", "/root/reference")
    if os.path.isdir(os.path.join(ref_root, "verl")):
        sys.dont_write_bytecode = True
        if ref_root not in sys.path:
            sys.path.insert(0, ref_root)
        try:
            from verl.utils.seqlen_balancing import get_seqlen_balanced_partitions as ref
        except Exception:  # tensordict is not installed here: load the two pure functions from the source file
            import types

            src = open(os.path.join(ref_root, "verl", "utils", "seqlen_balancing.py")).read()
            src = src.replace("from tensordict import TensorDict", "TensorDict = object")
            mod = types.ModuleType("_ref_seqlen_balancing")
            exec(compile(src, "seqlen_balancing.py", "exec"), mod.__dict__)
            ref = mod.get_seqlen_balanced_partitions
    for lens, cap in cases:
        parts = rearrange_micro_batches(lens, cap)
        num = -(-sum(lens) // cap)
        assert len(parts) == num and sorted(i for p in parts for i in p) == list(range(len(lens)))
        assert all(p == sorted(p) for p in parts)
        if ref is not None:
            assert parts == ref(lens, num, equal_size=False)
    with pytest.raises(AssertionError):
        rearrange_micro_batches([100, 5000], 4096)  # a sequence longer than max_token_len (seqlen_balancing.py:228)
    # TensorBatch.take builds the micro-batch the reference concatenates row by row (:245-251)
    tb = TensorBatch({"x": torch.arange(12).view(6, 2)}, {"uid": np.array(list("abcdef"), dtype=object)}, {"t": 1.0})
    mb = tb.take([4, 0, 5])
    assert mb.batch["x"].tolist() == [[8, 9], [0, 1], [10, 11]] and mb.non_tensor_batch["uid"].tolist() == ["e", "a", "f"]
    assert len(mb) == 3 and mb.meta_info == {"t": 1.0}


def test_balanced_rank_order_balances_every_optimizer_step():
    """sharding.balanced_rank_order: with mini_batches = 1 it is the reference's _balance_batch order
    (ray_trainer.py:526-541); with mini_batches = M every (rank, mini-batch) cell has the same number of sequences and
    nearly the same number of tokens, so the ranks reach each optimizer step's all-reduce together."""
    import random

    from spatialthinker_b200.sharding import balanced_partitions, balanced_rank_order

    rng = random.Random(11)
    world, minis, n = 8, 4, 1024
    lens = [rng.randint(1, 4096) for _ in range(n)]
    ref_order = [i for p in balanced_partitions(lens, world, equal_size=True) for i in p]
    assert balanced_rank_order(lens, world, 1) == ref_order
    local, mini = n // world, n // world // minis

    def cell_sums(order):
        return [[sum(lens[i] for i in order[r * local + m * mini: r * local + (m + 1) * mini]) for r in range(world)]
                for m in range(minis)]

    order = balanced_rank_order(lens, world, minis)
    assert sorted(order) == list(range(n))
    worst = max((max(c) - min(c)) / max(c) for c in cell_sums(order))
    worst_ref = max((max(c) - min(c)) / max(c) for c in cell_sums(ref_order))
    assert worst < 1e-3 < worst_ref  # per optimizer step: balanced to a few tokens vs several percent
    ranks = [sum(lens[i] for i in order[r * local:(r + 1) * local]) for r in range(world)]
    assert (max(ranks) - min(ranks)) / max(ranks) < 1e-3
    with pytest.raises(AssertionError):
        balanced_rank_order(lens[:100], 8, 4)


def test_speed_weighted_counts_and_weighted_cells():
    """Speed-aware shards: sequence counts proportional to measured speed (multiples of the optimizer steps, bounded shift,
    exact total) and, for ragged batches, cells with fixed sizes whose token sums follow the speeds."""
    import random

    from spatialthinker_b200.sharding import speed_weighted_counts, weighted_balanced_cells

    times = [1232.1, 1221.7, 1293.4, 1240.4, 1213.2, 1269.2, 1231.1, 1256.2]  # measured on one 8 x B200 box, equal shards
    counts = speed_weighted_counts(4096, times, 4)
    assert sum(counts) == 4096 and all(c % 4 == 0 for c in counts)
    pred = [c * t for c, t in zip(counts, times)]  # damped (x 0.75): a quarter of the 6.2 % spread is left on purpose
    assert (max(pred) - min(pred)) / max(pred) < 0.025 < (max(times) - min(times)) / max(times)
    full = [c * t for c, t in zip(speed_weighted_counts(4096, times, 4, damping=1.0), times)]
    assert (max(full) - min(full)) / max(full) < 0.012  # undamped: level to within one 4-sequence step
    assert speed_weighted_counts(4096, [1.0] * 8, 4) == [512] * 8
    wild = speed_weighted_counts(4096, [1.0, 1.0, 1.0, 5.0], 4, max_shift=0.10)  # a straggler cannot push anyone past +-10 %
    assert sum(wild) == 4096 and min(wild) >= 0.85 * 1024 and max(wild) <= 1.15 * 1024
    rng = random.Random(3)
    lens = [rng.randint(1, 4096) for _ in range(2048)]
    sizes = [c // 4 for c in speed_weighted_counts(2048, times, 4) for _ in range(4)]
    weights = [1.0 / t for t in times for _ in range(4)]
    cells = weighted_balanced_cells(lens, sizes, weights)
    assert [len(c) for c in cells] == sizes and sorted(i for c in cells for i in c) == list(range(2048))
    cost = [sum(lens[i] for i in c) / w for c, w in zip(cells, weights)]  # tokens x time per token
    assert (max(cost) - min(cost)) / max(cost) < 2e-3


def test_bench_global_layout_speed_aware():
    """bench.global_layout: equal shards reproduce the reference's balanced dispatch; given counts and step times, rank
    offsets follow the counts and every optimizer step's predicted time is level across the ranks."""
    import bench

    cfg = ("h", "v", 1024, 512, 8, "test")
    cfg = (64, 128) + cfg[2:]
    lens, uid, per_rank, naive, spread, counts, offsets = bench.global_layout(cfg, 4, True, 2)
    assert counts == [256] * 4 and offsets == [0, 256, 512, 768] and spread < 1e-3
    assert sorted(uid.tolist()) == sorted(list(range(128)) * 8)
    times = [100.0, 104.0, 97.0, 101.0]
    from spatialthinker_b200.sharding import speed_weights

    lens2, uid2, per2, _, spread2, counts2, offsets2 = bench.global_layout(cfg, 4, True, 2, [256] * 4, times)
    assert counts2 == [256] * 4 and spread2 < 2e-3  # equal sequence counts, tokens follow the (damped) speeds
    sw = speed_weights(times)
    assert sw[2] > sw[0] > sw[3] > sw[1] and max(sw) / min(sw) < 104.0 / 97.0  # damped
    share = [p / w for p, w in zip(per2, sw)]
    assert (max(share) - min(share)) / max(share) < 2e-3 and per2[2] > per2[1]
    assert sorted(lens2.tolist()) == sorted(lens.tolist()) and sorted(uid2.tolist()) == sorted(uid.tolist())
    # dense: nothing to balance but the counts
    _, _, per3, _, spread3, counts3, offsets3 = bench.global_layout(cfg, 4, False, 2, [252, 244, 268, 260], times)
    assert per3 == [c * 512 for c in [252, 244, 268, 260]] and offsets3 == [0, 252, 496, 764]


def test_actor_micro_plan_host_side():
    """DataParallelPPOActor._micro_plan (pure host logic): the reference's fixed-size split with GA (dp_actor.py:233-237);
    token-balanced micro-batches weighted by their share of the mini-batch; consecutive runs when all lengths are equal;
    a speed-aware shard's shorter tail micro-batch."""
    from spatialthinker_b200.dp_actor import ActorConfig, DataParallelPPOActor
    from spatialthinker_b200.sharding import rearrange_micro_batches

    w = torch.zeros(8, 8, dtype=torch.bfloat16)
    rows = list(range(100, 116))
    fixed = DataParallelPPOActor(ActorConfig(global_batch_size_per_device=16, micro_batch_size_per_device_for_update=4), w)
    plan = fixed._micro_plan(rows, None, 32)
    assert [p[0] for p in plan] == [rows[i:i + 4] for i in range(0, 16, 4)] and all(p[1] == 4.0 and p[2] is None for p in plan)
    with pytest.raises(AssertionError):
        fixed._micro_plan(rows[:15], None, 32)  # the reference's split() only knows equal chunks
    tail = DataParallelPPOActor(ActorConfig(global_batch_size_per_device=15, loss_scale_batch_size=16,
                                            micro_batch_size_per_device_for_update=4), w)._micro_plan(rows[:15], None, 32)
    assert [len(p[0]) for p in tail] == [4, 4, 4, 3] and [p[1] for p in tail] == [4.0, 4.0, 4.0, 16 / 3]
    dyn = DataParallelPPOActor(ActorConfig(global_batch_size_per_device=16, use_dynamic_bsz=True,
                                           max_token_len_per_micro_batch=100), w)
    dense = dyn._micro_plan(rows, [32] * 200, 32)  # lens are indexed by batch row
    assert [p[0] for p in dense] == [rows[0:3], rows[3:6], rows[6:9], rows[9:12], rows[12:14], rows[14:16]]  # ceil(512 / 100) = 6 runs
    assert all(abs(p[1] - 16 / len(p[0])) < 1e-12 and p[2] == 32 * len(p[0]) for p in dense)
    lens = {r: 5 + (7 * r) % 29 for r in rows}
    lens_list = [lens.get(i, 0) for i in range(200)]
    ragged = dyn._micro_plan(rows, lens_list, 32)
    want = rearrange_micro_batches([lens[r] for r in rows], 100)
    assert [p[0] for p in ragged] == [[rows[i] for i in part] for part in want]
    assert all(p[2] == sum(lens[r] for r in p[0]) for p in ragged)
