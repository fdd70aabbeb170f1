"""The peer-mapped dW exchange on N real GPUs, against NCCL: correctness at the 7B head's dW size and the time of the
optimizer-step exchange done both ways.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
        tools/gpu_peer_check.py [--rows 151936] [--hidden 3584] [--iters 10]

(a) NCCL:  all_reduce(AVG, fp32) -> grad_sumsq -> clip coefficient -> grad_scale_cast (bf16, zero)      [round-1 path]
(b) peer:  barrier -> reduce-scatter + sumsq -> barrier -> clip coefficient -> scale/cast/all-gather/zero -> barrier
Checks: (b)'s bf16 gradient equals bf16(clip * rank-ordered fp32 mean) bit for bit on sampled windows (the windows of all
ranks are all-gathered and summed in rank order on the device), is identical on every rank (checksums), and agrees with
(a)'s gradient to one bf16 rounding; the norms agree to 1e-6. Times are CUDA events per rank, max over ranks.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=151936)
    ap.add_argument("--hidden", type=int, default=3584)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--save", default=os.environ.get("GRPO_PEER_LOG", ""), help="file for rank 0's lines")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from spatialthinker_b200 import peer
    from spatialthinker_b200.dp_actor import grad_scale_cast, grad_sumsq

    lines = []

    def say(msg):
        if rank == 0:
            print(msg, flush=True)
            lines.append(msg)

    n = args.rows * args.hidden
    max_norm = 1.0
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    source = torch.randn(n, generator=gen, device=dev) * (1.0 + 0.25 * rank) * 1e-3
    group = peer.get_group(None, dev)
    grad = source.clone()
    out_peer = torch.zeros(n, dtype=torch.bfloat16, device=dev)
    gbuf, obuf = group.register(grad), group.register(out_peer)
    say(f"peer group up: {world} ranks, dW {args.rows} x {args.hidden} fp32 = {n * 4 / 1e9:.2f} GB per rank")

    def nccl_path(g, out):
        dist.all_reduce(g, op=dist.ReduceOp.AVG)
        norm = grad_sumsq(g).sqrt().float()
        clip = torch.clamp(max_norm / (norm + 1e-6), max=1.0).reshape(1)
        grad_scale_cast(g, clip, out, zero_after=True)
        return norm

    def peer_path():
        norm = group.reduce_scatter_sumsq(gbuf).sqrt().float()
        clip = torch.clamp(max_norm / (norm + 1e-6), max=1.0).reshape(1)
        group.scale_cast_allgather(gbuf, obuf, clip, zero_after=True)
        return norm, clip

    # ---- correctness
    norm_p, clip_p = peer_path()
    g2 = source.clone()
    out_nccl = torch.zeros(n, dtype=torch.bfloat16, device=dev)
    norm_n = nccl_path(g2, out_nccl)
    torch.cuda.synchronize()
    ok = True
    rel = abs(float(norm_p) - float(norm_n)) / float(norm_n)
    ok &= rel <= 1e-6
    ok &= not bool(grad.any()) and not bool(g2.any())
    diff = (out_peer.float() - out_nccl.float()).abs().max()
    scale = out_nccl.float().abs().max()
    ok &= float(diff) <= 2 ** -7 * float(scale)
    # identical bits on every rank
    check = out_peer.view(torch.int16).to(torch.int64).sum().reshape(1)
    lo, hi = check.clone(), check.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok &= int(lo) == int(hi)
    # rank-ordered reference on windows: start, every slab edge, end
    win = 1 << 16
    starts = {0, n - win}
    for r in range(world):
        e0, e1 = peer.slab_bounds(n, r, world)
        starts |= {max(0, e0 - win // 2), max(0, min(n - win, e1 - win // 2))}
    exact = True
    for s0 in sorted(starts):
        mine = source[s0:s0 + win].contiguous()
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        acc = every[0].clone()
        for e in every[1:]:
            acc += e
        acc *= torch.tensor(1.0, dtype=torch.float32, device=dev) / world
        want = (acc * clip_p).to(torch.bfloat16)
        exact &= torch.equal(want.view(torch.int16), out_peer[s0:s0 + win].view(torch.int16))
    ok &= exact
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(f"[rank {rank}] norm peer {float(norm_p):.9g} nccl {float(norm_n):.9g} (rel {rel:.1e})  max|peer - nccl| "
          f"{float(diff):.3e} of {float(scale):.3e}  windows bit-exact {exact}  checksum {int(check)}", flush=True)
    dist.barrier()
    say("PEER CHECK " + ("OK" if float(flag) == 1.0 else "FAILED"))

    # ---- timing (events per rank, max over ranks; every iteration starts from a filled accumulator)
    def timed(fn, prepare):
        times = []
        for it in range(args.iters + 2):
            prepare()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it >= 2:
                times.append(float(t))
        times.sort()
        return times[len(times) // 2], times[0], times[-1]

    wire_ar = 2 * (world - 1) / world * n * 4
    wire_peer = (world - 1) / world * n * (4 + 2)
    t = timed(lambda: nccl_path(g2, out_nccl), lambda: g2.copy_(source))
    say(f"NCCL all-reduce + norm + clip/cast/zero : median {t[0]:7.3f} ms  (min {t[1]:.3f} max {t[2]:.3f})  "
        f"wire {wire_ar / 1e9:.2f} GB per direction and GPU")
    t_ar = timed(lambda: dist.all_reduce(g2, op=dist.ReduceOp.AVG), lambda: g2.copy_(source))
    say(f"  of which ncclAllReduce alone          : median {t_ar[0]:7.3f} ms  -> {wire_ar / t_ar[0] / 1e6:.0f} GB/s per direction")
    tp = timed(lambda: peer_path(), lambda: grad.copy_(source))
    say(f"peer reduce-scatter+norm, clip+bf16 all-gather+zero : median {tp[0]:7.3f} ms  (min {tp[1]:.3f} max {tp[2]:.3f})  "
        f"wire {wire_peer / 1e9:.2f} GB -> {wire_peer / tp[0] / 1e6:.0f} GB/s per direction")
    t_rs = timed(lambda: group.reduce_scatter_sumsq(gbuf), lambda: grad.copy_(source))
    say(f"  of which reduce-scatter + sumsq (2 barriers)      : median {t_rs[0]:7.3f} ms  -> "
        f"{(world - 1) / world * n * 4 / t_rs[0] / 1e6:.0f} GB/s per direction")
    t_pa = timed(lambda: group.allreduce_mean_(gbuf), lambda: grad.copy_(source))
    say(f"peer in-place fp32 all-reduce (for comparison)      : median {t_pa[0]:7.3f} ms  -> {wire_ar / t_pa[0] / 1e6:.0f} GB/s per direction")
    say(f"optimizer-step exchange: {t[0] / tp[0]:.2f}x faster than the NCCL path")
    group.release(gbuf)
    group.release(obuf)
    if rank == 0 and args.save:
        with open(args.save, "w") as f:
            f.write("\n".join(lines) + "\n")
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
