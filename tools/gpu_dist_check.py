"""Numerical check of the N > 1 path on real GPUs (NCCL), the GPU twin of tests/test_distributed_cpu.py:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/gpu_dist_check.py

Every rank builds the same global rollout (seeded), takes its rows, and runs the product path: sharded advantages (score
all-gather over NCCL, group statistics per rank), one fused micro-batch step on its shard, mean all-reduce of dW. The
results are compared with (a) the CPU oracle's advantages and (b) the same steps done for every shard on ONE GPU and
averaged locally - so the only thing under test is the exchange itself.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import spatialthinker_b200 as st
    from oracle import grpo_oracle as O
    from spatialthinker_b200.sharding import allreduce_mean_

    bsz, t, h, v, n = 8 * world, 64, 256, 8192, 4
    roll = O.synth_rollout(bsz, t, v, n, seed=3, ragged=True)  # identical on every rank; groups straddle the ranks
    hidden, weight = O.synth_head(bsz * t, h, v, seed=4, sigma_w=0.05)
    hidden = hidden.view(bsz, t, h)
    logp, _ = O.lm_head_log_probs(hidden, weight, roll["responses"])
    old, ref = O.perturbed_log_probs(logp, seed=5), O.perturbed_log_probs(logp, seed=6)
    want_adv, _ = O.compute_grpo_outcome_advantage(roll["token_level_rewards"].clone(), roll["response_mask"], roll["uid"])
    per = bsz // world
    w_d = weight.to(dev)

    def shard_step(r, adv_rows):
        sl = slice(r * per, (r + 1) * per)
        dw = torch.zeros(v, h, dtype=torch.float32, device=dev)
        res = st.grpo_micro_batch_step(hidden[sl].to(dev), w_d, roll["responses"][sl].to(dev), old[sl].to(dev), adv_rows,
                                       ref[sl].to(dev), roll["response_mask"][sl].to(dev), kl_penalty="low_var_kl",
                                       kl_coef=1e-2, grad_accum=2.0, dweight_accum=dw)
        return dw, res["metrics"]

    sl = slice(rank * per, (rank + 1) * per)
    adv, ret = st.core_algos.compute_grpo_outcome_advantage_sharded(
        roll["token_level_rewards"][sl].to(dev), roll["response_mask"][sl].to(dev), roll["uid"], rank * per)
    assert adv is ret
    err_adv = float((adv.cpu() - want_adv[sl]).abs().max())
    dw, met = shard_step(rank, adv)
    allreduce_mean_(dw)
    # the same on one GPU: every shard in turn, averaged locally
    want_dw = torch.zeros_like(dw)
    for r in range(world):
        d, _ = shard_step(r, want_adv[r * per:(r + 1) * per].to(dev))
        want_dw += d
    want_dw /= world
    err_dw = float((dw - want_dw).norm() / want_dw.norm())
    ok = torch.tensor([1.0 if (err_adv <= 1e-6 and err_dw <= 1e-5) else 0.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    print(f"[rank {rank}/{world}] advantages max|err| {err_adv:.2e}  dW rel err vs local average {err_dw:.2e}  "
          f"loss {float(met[7]):+.6f}", flush=True)
    dist.barrier()
    if rank == 0:
        print("DIST CHECK", "OK" if float(ok) == 1.0 else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if float(ok) == 1.0 else 1)


if __name__ == "__main__":
    main()
