"""Peer-mapped dW exchange (csrc/peer_kernels.cuh, spatialthinker_b200/peer.py): reduce-scatter + norm, clip + bf16
all-gather + zero - what replaces FSDP's fp32 gradient averaging (verl/workers/actor/config.py:58) and the passes of
``_optimizer_step`` (verl/workers/actor/dp_actor.py:155-167) for the replicated lm_head weight.

CPU: the slab partition and the argument checks of the C ABI. GPU, one process: the kernels with the W "ranks" played
by W buffers of one device (no barrier needed: one stream). GPU, two processes on ONE device: the real thing - CUDA IPC
mappings, flag barriers, the actor's optimizer step - against the collective path (gloo here; NCCL on two GPUs:
tools/gpu_peer_check.py)."""
import ctypes
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_partition_covers_every_element_once():
    from spatialthinker_b200.peer import slab_bounds

    for n in (8, 64, 8 * 1000, 151936 * 3584, 8 * 7):
        for world in range(1, 9):
            edges = [slab_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0 and a0 <= a1 and b0 <= b1
            assert all(e0 % 8 == 0 and e1 % 8 == 0 for e0, e1 in edges)
            assert max(e1 - e0 for e0, e1 in edges) <= 8 * (-(-(n // 8) // world))


def test_host_mirror_of_the_slab_partition_matches_the_library():
    """peer.slab_bounds (used by tests and tools to know which rank reduces what) against the partition the library
    launches its kernels with (grpo_debug_peer_slab, host-only)."""
    from spatialthinker_b200 import _lib
    from spatialthinker_b200.peer import slab_bounds

    lib = _lib.load()
    e0, e1 = ctypes.c_int64(-1), ctypes.c_int64(-1)
    for n in (0, 8, 8 * 7, 8 * 1001, 151936 * 3584, 152064 * 2048):
        for world in range(1, 9):
            for rank in range(world):
                assert lib.grpo_debug_peer_slab(n, rank, world, ctypes.byref(e0), ctypes.byref(e1)) == 0
                assert (e0.value, e1.value) == slab_bounds(n, rank, world), (n, rank, world)
    assert lib.grpo_debug_peer_slab(12, 0, 2, ctypes.byref(e0), ctypes.byref(e1)) == -1
    assert lib.grpo_debug_peer_slab(16, 2, 2, ctypes.byref(e0), ctypes.byref(e1)) == -1


def test_peer_entry_points_reject_bad_arguments():
    """Argument checks come before any CUDA call: they answer GRPO_ERR_ARG on a box without a GPU too."""
    from spatialthinker_b200 import _lib

    lib = _lib.load()
    table = (ctypes.c_void_p * 8)(*([4096] * 8))
    null_table = (ctypes.c_void_p * 8)()
    assert lib.grpo_peer_barrier(table, 0, 0, 1, 0, None) == -1          # world < 1
    assert lib.grpo_peer_barrier(table, 0, 9, 1, 0, None) == -1          # world > GRPO_MAX_PEERS
    assert lib.grpo_peer_barrier(table, 3, 2, 1, 0, None) == -1          # rank outside the world
    assert lib.grpo_peer_barrier(null_table, 0, 2, 1, 0, None) == -1     # null flag pointer
    assert lib.grpo_peer_allreduce_mean(table, 0, 2, 6, None) == -1      # n % 4
    assert lib.grpo_peer_reduce_scatter_sumsq(table, table, 0, 2, 12, ctypes.c_void_p(4096), None) == -1   # n % 8
    assert lib.grpo_peer_reduce_scatter_sumsq(table, table, 0, 2, 16, None, None) == -1                    # no scratch
    odd = (ctypes.c_void_p * 8)(*([4100] * 8))
    assert lib.grpo_peer_reduce_scatter_sumsq(odd, table, 0, 2, 16, ctypes.c_void_p(4096), None) == -1     # alignment
    assert lib.grpo_peer_scale_cast_allgather(None, table, 0, 2, 16, None, 1.0, 1, None) == -1
    assert lib.grpo_peer_scale_cast_allgather(ctypes.c_void_p(4096), table, 0, 2, 20, None, 1.0, 1, None) == -1
    assert b"multiple of 8" in lib.grpo_last_error()
    assert lib.grpo_ipc_export(None, None, None) == -1
    assert lib.grpo_ipc_open(None, 0, None) == -1
    assert lib.grpo_ipc_close(None, 0) == 0


def _expected(copies, max_norm):
    """fp32 mean in rank order, global norm in fp64, clip coefficient, bf16 gradient - the CPU oracle's restatement."""
    from oracle import grpo_oracle as O

    want = O.averaged_clipped_gradient(copies, max_norm)
    return want["mean"], want["norm"], want["clip"], want["grad"]


def test_oracle_of_the_exchange_is_torch_clip_grad_norm_on_the_fp32_mean():
    """Pins oracle.averaged_clipped_gradient to the functions the reference itself calls (dp_actor.py:155-167):
    ``clip_grad_norm_`` over several parameters holding the rank-averaged fp32 gradients."""
    from oracle import grpo_oracle as O

    gen = torch.Generator().manual_seed(3)
    for world, max_norm in ((2, 0.5), (3, 1.0), (8, 1e6)):
        grads = [torch.randn(4096, generator=gen) * (1 + q) for q in range(world)]
        other = torch.randn(777, generator=gen)
        head = torch.nn.Parameter(torch.zeros(4096))
        body = torch.nn.Parameter(torch.zeros(777))
        head.grad = torch.stack(grads).sum(0) / world
        body.grad = other.clone()
        total = torch.nn.utils.clip_grad_norm_([head, body], max_norm=max_norm)
        got = O.averaged_clipped_gradient(grads, max_norm, other_sumsq=float(other.double().square().sum()))
        assert abs(float(got["norm"]) - float(total)) <= 1e-5 * float(total)
        assert torch.allclose(got["mean"] * got["clip"], head.grad, rtol=1e-5, atol=1e-8)
        assert got["grad"].dtype == torch.bfloat16 and (float(got["clip"]) == 1.0) == (max_norm == 1e6)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("n", [8 * 5, 8 * 4099, 8 * 148 * 4 * 512 * 2 + 8 * 77])
def test_peer_kernels_with_ranks_played_by_local_buffers(world, n):
    from spatialthinker_b200 import _lib
    from spatialthinker_b200.peer import slab_bounds

    lib = _lib.load()
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(1000 * world + n % 997)
    host = [torch.randn(n, generator=gen) * (1.0 + q) for q in range(world)]
    copies = [h.to(dev) for h in host]
    outs = [torch.full((n,), 7.0, dtype=torch.bfloat16, device=dev) for _ in range(world)]
    partials = [torch.zeros(8, dtype=torch.float64, device=dev) for _ in range(world)]
    scratch = torch.zeros(_lib.GRAD_SCRATCH_DOUBLES, dtype=torch.float64, device=dev)

    def table(ts):
        return (ctypes.c_void_p * 8)(*([t.data_ptr() for t in ts] + [0] * (8 - len(ts))))

    stream = _lib.stream_ptr(dev)
    for r in range(world):
        _lib.check(lib.grpo_peer_reduce_scatter_sumsq(table(copies), table(partials), r, world, n, scratch.data_ptr(), stream),
                   "reduce_scatter")
    want, norm, clip, want_bf16 = _expected(host, max_norm=1.0)
    for r in range(world):
        e0, e1 = slab_bounds(n, r, world)
        assert torch.equal(copies[r][e0:e1].cpu(), want[e0:e1]), f"slab of rank {r}"          # bit-exact: same order
        for q in range(world):                                                               # the rest is untouched
            if q != r:
                assert torch.equal(copies[q][e0:e1].cpu(), host[q][e0:e1])
    assert int(scratch.view(torch.int64)[148 * 4].item()) == 0  # the ticket is left at zero for the next launch
    for q in range(1, world):
        assert torch.equal(partials[q], partials[0])
    got_norm = partials[0][:world].sum().sqrt()
    assert abs(float(got_norm) - float(norm)) <= 1e-6 * float(norm)
    assert float(clip) < 1.0
    clip_dev = clip.reshape(1).to(dev)
    for r in range(world):
        _lib.check(lib.grpo_peer_scale_cast_allgather(copies[r].data_ptr(), table(outs), r, world, n, clip_dev.data_ptr(),
                                                      1.0, 1, stream), "scale_cast_allgather")
    for q in range(world):
        assert torch.equal(outs[q].cpu().view(torch.int16), want_bf16.view(torch.int16)), f"gradient buffer of rank {q}"
        assert not copies[q].any(), f"accumulator of rank {q} not zeroed"
    # general in-place all-reduce (n % 4)
    copies = [h.to(dev) for h in host]
    for r in range(world):
        _lib.check(lib.grpo_peer_allreduce_mean(table(copies), r, world, n, stream), "allreduce_mean")
    for q in range(world):
        assert torch.equal(copies[q].cpu(), want)


# ----------------------------------------------------------------------------------------------------------------
# two processes, one GPU: IPC + barriers + the actor
# ----------------------------------------------------------------------------------------------------------------
def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      GRPO_PEER_TIMEOUT_MS="60000")
    import torch.distributed as dist

    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import spatialthinker_b200 as st
        from oracle import grpo_oracle as O
        from spatialthinker_b200 import peer
        from spatialthinker_b200.dp_actor import ActorConfig, DataParallelPPOActor
        from spatialthinker_b200.protocol import TensorBatch

        dev = torch.device("cuda:0")
        # (1) the exchange itself, twice (epochs advance, the ticket resets)
        group = peer.get_group(None, dev)
        n = 8 * 30011
        host = [torch.randn(n, generator=torch.Generator().manual_seed(50 + q)) * (1 + q) for q in range(world)]
        grad = host[rank].to(dev)
        out_bf16 = torch.zeros(n, dtype=torch.bfloat16, device=dev)
        gbuf, obuf = group.register(grad), group.register(out_bf16)
        for rep in range(2):
            grad.copy_(host[rank])
            want, norm, clip, want_bf16 = _expected(host, max_norm=1.0)
            total = group.reduce_scatter_sumsq(gbuf)
            got_norm = total.sqrt()
            group.scale_cast_allgather(gbuf, obuf, clip.reshape(1).to(dev), zero_after=True)
            assert abs(float(got_norm) - float(norm)) <= 1e-6 * float(norm)
            assert torch.equal(out_bf16.cpu().view(torch.int16), want_bf16.view(torch.int16))
            assert not grad.any()
        grad.copy_(host[rank])
        group.allreduce_mean_(gbuf)
        assert torch.equal(grad.cpu(), want)
        group.release(gbuf)
        group.release(obuf)

        # (2) the actor: the same update through the peer exchange and through the collective
        bsz, t, h, v, ng = 8 * world, 32, 256, 8192, 4
        roll = O.synth_rollout(bsz, t, v, ng, seed=3, ragged=True)
        hidden, weight = O.synth_head(bsz * t, h, v, seed=4, sigma_w=0.05)
        hidden = hidden.view(bsz, t, h)
        logp, _ = O.lm_head_log_probs(hidden, weight, roll["responses"])
        old = O.perturbed_log_probs(logp, seed=5)
        adv, _ = O.compute_grpo_outcome_advantage(roll["token_level_rewards"].clone(), roll["response_mask"], roll["uid"])
        per = bsz // world
        sl = slice(rank * per, (rank + 1) * per)
        results = {}
        for mode in (True, False):
            w = weight.to(dev).to(torch.bfloat16).requires_grad_(True)
            opt = torch.optim.SGD([w], lr=0.5)
            cfg = ActorConfig(global_batch_size_per_device=per // 2, micro_batch_size_per_device_for_update=2,
                              max_grad_norm=0.05, clip_ratio_low=0.2, clip_ratio_high=0.3, clip_ratio_dual=3.0)
            actor = DataParallelPPOActor(cfg, w, actor_optimizer=opt, peer_exchange=mode, defer_dw=False)
            batch = TensorBatch({"hidden_states": hidden[sl].to(dev).to(torch.bfloat16), "responses": roll["responses"][sl].to(dev),
                                 "old_log_probs": old[sl].to(dev), "advantages": adv[sl].to(dev),
                                 "response_mask": roll["response_mask"][sl].to(dev)}, meta_info={"temperature": 1.0})
            metrics = actor.update_policy(batch)
            assert bool(actor._peer) == mode
            results[mode] = (metrics["actor/grad_norm"], w.detach().float().cpu().clone(), metrics["actor/pg_loss"])
            actor.release_workspaces()
        (norm_p, w_p, loss_p), (norm_c, w_c, loss_c) = results[True], results[False]
        assert len(norm_p) == 2 and all(abs(a - b) <= 1e-5 * abs(b) for a, b in zip(norm_p, norm_c)), (norm_p, norm_c)
        assert all(g > 0.05 for g in norm_c)  # the clip coefficient was in play
        # the weights moved, and moved alike: bf16 gradients differ by at most one rounding of the last bit
        assert (w_c - weight.to(torch.bfloat16).float()).abs().max() > 0
        assert (w_p - w_c).abs().max() <= 2 ** -7 * (w_c - weight.to(torch.bfloat16).float()).abs().max() + 1e-6
        assert loss_p[:per // 4] == loss_c[:per // 4]  # first optimizer step: identical weights, identical losses
        # every rank ends with the same weights
        gathered = [None] * world
        dist.all_gather_object(gathered, w_p)
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_processes_share_one_gpu_through_ipc():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"
