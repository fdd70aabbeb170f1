"""Sharding of the GRPO hot path over the GPUs of one NVSwitch box: one process per GPU, sequences are the unit.

* The ``lm_head`` weight is replicated; each rank owns an equal number of rollout sequences, token-balanced with the
  largest-differencing method - the same policy the reference's driver applies before dispatch
  (``_balance_batch`` ray_trainer.py:526-541 -> ``get_seqlen_balanced_partitions`` seqlen_balancing.py:150-181).
* Advantages need every sequence's score (groups straddle ranks after balancing): scores are all-gathered
  (``B`` floats) and the group statistics run redundantly per rank.
* The only data-path collective is the mean all-reduce of ``dW_lm_head`` (fp32, once per optimizer step), matching
  FSDP's gradient averaging with ``mp_reduce_dtype = fp32`` (actor/config.py:58).
"""
from __future__ import annotations

import heapq
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


# ----------------------------------------------------------------------------------------------------------------
# token-balanced partition of sequences
# ----------------------------------------------------------------------------------------------------------------
def balanced_partitions(seqlens: Sequence[int], k: int, equal_size: bool = True) -> List[List[int]]:
    """Split item indices into ``k`` parts with near-equal length sums (Karmarkar-Karp largest differencing).

    ``equal_size=True`` additionally forces the same number of items per part (``len(seqlens) % k == 0``), which is
    what the dispatch to data-parallel ranks needs. Each returned part is sorted by index, like
    seqlen_balancing.py:165-176.
    """
    n = len(seqlens)
    assert n >= k, f"number of items:[{n}] < k_partitions:[{k}]"
    # A partial solution ("state") is a list of k bins kept in DESCENDING order; a bin is the tuple
    # (sum, count, items) with items a tuple of (index, length) pairs - exactly the fields the reference's Set.__lt__
    # compares, in its order (seqlen_balancing.py:39-44), so plain tuple comparison reproduces its ordering without a key
    # function. Two states are merged by pairing the heaviest bin of one with the lightest of the other. The heap pops
    # the state with the largest spread (heaviest - lightest) first, ties going to the state whose heaviest bin is
    # larger (State.__lt__, :77-83).
    empty = (0, 0, ())

    def make_state(items):
        bins = [(length, 1, ((idx, length),)) for idx, length in items] + [empty] * (k - len(items))
        bins.sort(reverse=True)
        return bins

    def heap_entry(bins, serial):
        # min-heap keys, all plain ints: -spread, then "the larger heaviest bin first" = descending (sum, count, items).
        # Item tuples of two different bins never share their first (index, length) pair (indices are distinct), so the
        # first index decides wherever sum and count tie. `serial` keeps comparisons total.
        top = bins[0]
        return (bins[-1][0] - top[0], -top[0], -top[1], -(top[2][0][0] if top[1] else -1), serial, bins)

    by_len = sorted((int(length), idx) for idx, length in enumerate(seqlens))
    heap = []
    serial = 0
    if equal_size:
        assert n % k == 0, f"{n} % {k} != 0"
        for off in range(0, n, k):
            items = [(idx, length) for length, idx in by_len[off:off + k]]
            heap.append(heap_entry(make_state(items), serial))
            serial += 1
    else:
        for length, idx in by_len:
            heap.append(heap_entry(make_state([(idx, length)]), serial))
            serial += 1
    heapq.heapify(heap)
    last = k - 1
    while len(heap) > 1:
        a = heapq.heappop(heap)[5]
        b = heapq.heappop(heap)[5]
        merged = []
        for i in range(k):
            x, y = a[i], b[last - i]
            merged.append((x[0] + y[0], x[1] + y[1], x[2] + y[2]) if y[1] else x)
        merged.sort(reverse=True)
        heapq.heappush(heap, heap_entry(merged, serial))
        serial += 1
    parts = [sorted(idx for idx, _ in items) for _, _, items in heap[0][5]]
    seen = sorted(i for p in parts for i in p)
    assert seen == list(range(n)) and all(len(p) > 0 for p in parts)
    if equal_size:
        assert all(len(p) * k == n for p in parts)
    return parts


def micro_batch_counts(token_sums: Sequence[int], max_token_len: int, group: Optional["dist.ProcessGroup"] = None,
                       device=None) -> List[int]:
    """``ceil(tokens / max_token_len)`` per mini-batch, raised to the maximum over the data-parallel ranks so that every
    rank runs the same number of micro-batches (seqlen_balancing.py:234-239). ONE collective for all mini-batches of a
    call instead of one per mini-batch."""
    nums = [max(1, -(-int(t) // int(max_token_len))) for t in token_sums]
    if nums and dist.is_available() and dist.is_initialized() and _world(group) > 1:
        backend = dist.get_backend(group)
        t = torch.tensor(nums, dtype=torch.int64, device=device if backend == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        nums = [int(v) for v in t.tolist()]
    return nums


def rearrange_micro_batches(seq_lens: Sequence[int], max_token_len: int, group: Optional["dist.ProcessGroup"] = None,
                            device=None, num_micro_batches: Optional[int] = None) -> List[List[int]]:
    """Index lists of token-balanced micro-batches: seqlen_balancing.py:222-255 on the host-known token counts.

    ``ceil(sum(seq_lens) / max_token_len)`` micro-batches (the maximum of that over the data-parallel ranks, :236-239 -
    or ``num_micro_batches`` when the caller has already agreed on it, see :func:`micro_batch_counts`), items assigned
    by :func:`balanced_partitions` with ``equal_size=False``, each list sorted by index. ``max_token_len`` must be at
    least the longest sequence (:228-230)."""
    lens = [int(x) for x in seq_lens]
    assert lens and max_token_len >= max(lens), (
        f"max_token_len must be greater than the sequence length. Got {max_token_len=} and max_seq_len={max(lens) if lens else 0}")
    num = num_micro_batches if num_micro_batches is not None else micro_batch_counts([sum(lens)], max_token_len, group, device)[0]
    assert num <= len(lens)
    return balanced_partitions(lens, num, equal_size=False)


def rank_rows(seqlens: Sequence[int], world_size: int, rank: int) -> List[int]:
    """Row ids this rank owns after token-balancing (equal sequence counts)."""
    return balanced_partitions(seqlens, world_size, equal_size=True)[rank]


def balanced_rank_order(seqlens: Sequence[int], world_size: int, mini_batches: int = 1) -> List[int]:
    """Row order for dispatching a rollout batch to ``world_size`` data-parallel ranks: rank ``r`` owns rows
    ``order[r * local : (r + 1) * local]`` (``local = len(seqlens) / world_size``).

    ``mini_batches == 1`` is the reference's ``_balance_batch`` (ray_trainer.py:526-541): one Karmarkar-Karp partition of
    the whole batch, equal sequence counts, balanced token sums per rank. The ranks meet at EVERY optimizer step (the
    gradient all-reduce), though, and the reference then cuts each rank's shard into consecutive mini-batches
    (dp_actor.py:227) whose token sums are whatever they happen to be - at 256 ragged sequences per mini-batch they differ
    by +-3.6 % between ranks and the fast ranks wait at each all-reduce. ``mini_batches = M > 1`` balances both ways: the
    batch is first split into ``M`` groups of equal size and token sum, then each group over the ranks; rank ``r``'s
    ``m``-th consecutive mini-batch is its part of group ``m``, so every (rank, mini-batch) cell carries the same number of
    sequences and (nearly) the same number of tokens."""
    n = len(seqlens)
    assert n % (world_size * mini_batches) == 0, f"{n} % ({world_size} * {mini_batches}) != 0"
    lens = [int(x) for x in seqlens]
    groups = balanced_partitions(lens, mini_batches, equal_size=True) if mini_batches > 1 else [list(range(n))]
    cells = []  # cells[m][r] = global row ids
    for group in groups:
        parts = balanced_partitions([lens[i] for i in group], world_size, equal_size=True)
        cells.append([[group[j] for j in part] for part in parts])
    return [i for r in range(world_size) for m in range(len(groups)) for i in cells[m][r]]


# ----------------------------------------------------------------------------------------------------------------
# speed-aware sharding: chips under a power cap do not run at the same clock
# ----------------------------------------------------------------------------------------------------------------
SPEED_DAMPING = 0.75


def speed_weights(step_times: Sequence[float], damping: float = SPEED_DAMPING, max_shift: float = 0.10) -> List[float]:
    """Relative share of the work each rank should get, from its measured time per step on EQUAL work.

    ``(mean_time / time) ** damping``, clipped to ``1 +- max_shift``. The damping is empirical: a chip that used to wait
    at the all-reduce 5 % of the time and is then kept busy loses ~0.3 % of speed per percent of idle time it gives up
    (power cap / temperature; profiles/README.md, round 2), so handing it the full ``1 / time`` share overshoots."""
    times = [max(float(t), 1e-9) for t in step_times]
    mean = sum(times) / len(times)
    return [min(max((mean / t) ** damping, 1.0 - max_shift), 1.0 + max_shift) for t in times]


def speed_weighted_counts(total: int, step_times: Sequence[float], multiple: int = 1, max_shift: float = 0.10,
                          damping: float = SPEED_DAMPING) -> List[int]:
    """Sequences per rank in proportion to :func:`speed_weights`, each a multiple of ``multiple``, summing to ``total``.

    Eight B200s of one box at the 1000 W cap differ by 4-7 % in sustained GEMM throughput (profiles/README.md, round 2), and
    every optimizer step ends in an all-reduce: with equal shards the fast chips idle for that difference. The reference
    assumes homogeneous GPUs (equal shards, ray_trainer.py:526-541). No rank's share moves by more than ``max_shift``."""
    world = len(step_times)
    assert total % multiple == 0 and total // multiple >= world
    speeds = speed_weights(step_times, damping, max_shift)
    units = total // multiple
    ideal = [units * v / sum(speeds) for v in speeds]
    counts = [max(1, int(x)) for x in ideal]
    # largest remainders get the units still to hand out (or give back)
    order = sorted(range(world), key=lambda r: ideal[r] - counts[r], reverse=True)
    i = 0
    while sum(counts) < units:
        counts[order[i % world]] += 1
        i += 1
    while sum(counts) > units:
        r = max(range(world), key=lambda q: counts[q] - ideal[q])
        counts[r] -= 1
    return [c * multiple for c in counts]


def weighted_balanced_cells(seqlens: Sequence[int], cell_sizes: Sequence[int], cell_weights: Sequence[float],
                            refine_iters: int = 400) -> List[List[int]]:
    """Assign every sequence to one of ``len(cell_sizes)`` cells with EXACTLY ``cell_sizes[c]`` sequences each and token
    sums proportional to ``cell_weights`` (a cell = one rank's mini-batch; weight = that rank's speed).

    Longest-processing-time greedy under capacity, then pairwise swaps between the most over- and under-filled cells.
    With equal sizes and weights use :func:`balanced_rank_order` (the reference's Karmarkar-Karp) instead."""
    import numpy as np

    lens = np.asarray([int(x) for x in seqlens], dtype=np.int64)
    sizes = [int(c) for c in cell_sizes]
    assert sum(sizes) == len(lens) and all(c > 0 for c in sizes)
    w = np.asarray(cell_weights, dtype=np.float64)
    target = lens.sum() * w / w.sum()
    cells: List[List[int]] = [[] for _ in sizes]
    sums = np.zeros(len(sizes))
    free = np.asarray(sizes, dtype=np.int64)
    for i in np.argsort(-lens, kind="stable"):
        # the cell that is furthest below its target PER FREE SLOT keeps long and short sequences mixed
        score = np.where(free > 0, (target - sums) / np.maximum(free, 1), -np.inf)
        c = int(np.argmax(score))
        cells[c].append(int(i))
        sums[c] += lens[i]
        free[c] -= 1
    for _ in range(refine_iters):
        dev = sums - target
        a, b = int(np.argmax(dev)), int(np.argmin(dev))
        if a == b or dev[a] - dev[b] < 2:
            break
        ia, ib = np.asarray(cells[a]), np.asarray(cells[b])
        want = (dev[a] - dev[b]) / 2.0  # moving `want` tokens from a to b equalises the two deviations
        diff = lens[ia][:, None] - lens[ib][None, :]
        gain = np.abs(diff - want)
        k = int(np.argmin(gain))
        x, y = divmod(k, len(ib))
        d = int(diff[x, y])
        if d <= 0 or d >= dev[a] - dev[b]:
            break  # no swap brings the pair closer
        cells[a][x], cells[b][y] = int(ib[y]), int(ia[x])
        sums[a] -= d
        sums[b] += d
    return [sorted(c) for c in cells]


# ----------------------------------------------------------------------------------------------------------------
# collectives
# ----------------------------------------------------------------------------------------------------------------
def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_mean_(grad: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """In-place mean all-reduce of the fp32 ``dW_lm_head`` accumulator over the data-parallel ranks."""
    ws = _world(group)
    if ws == 1:
        return grad
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(grad, op=dist.ReduceOp.AVG, group=group)
    else:  # gloo (CPU tests) has no AVG
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
        grad.div_(ws)
    return grad


def all_gather_rows(local: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                    sizes: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Concatenate per-rank row blocks (sequence scores, uid codes) in rank order. ``sizes`` (rows per rank, known to
    every rank) allows unequal blocks (speed-aware shards): blocks are padded to the largest for the collective."""
    ws = _world(group)
    if ws == 1:
        return local
    if sizes is None:
        out = [torch.empty_like(local) for _ in range(ws)]
        dist.all_gather(out, local.contiguous(), group=group)
        return torch.cat(out, dim=0)
    assert len(sizes) == ws and local.shape[0] == sizes[dist.get_rank(group)]
    cap = max(sizes)
    padded = local.new_zeros((cap,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    out = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)
