"""world_size-2 gloo tests of the N>1 host path: sequence sharding, score all-gather, mean all-reduce of dW."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import grpo_oracle as O
        from spatialthinker_b200.sharding import all_gather_rows, allreduce_mean_, rank_rows, rearrange_micro_batches

        bsz, t, v, hd, n = 16, 6, 128, 64, 4
        roll = O.synth_rollout(bsz, t, v, n, seed=3, ragged=True)  # identical on every rank
        hidden, weight = O.synth_head(bsz * t, hd, v, seed=4)
        hidden = hidden.view(bsz, t, hd)
        lens = roll["response_mask"].sum(-1).tolist()
        mine = rank_rows(lens, world, rank)
        # (1) every rank scores its own sequences; the all-gather restores the global score vector
        local_scores = roll["token_level_rewards"][mine].sum(-1)
        gathered = all_gather_rows(local_scores)
        owners = [rank_rows(lens, world, r) for r in range(world)]
        full = torch.empty(bsz)
        full[torch.tensor([i for o in owners for i in o])] = gathered
        assert torch.equal(full, roll["token_level_rewards"].sum(-1))
        # (2) advantages are computed from the global scores (groups straddle ranks), then sliced
        adv, _ = O.compute_grpo_outcome_advantage(roll["token_level_rewards"].clone(), roll["response_mask"], roll["uid"])
        logp, _ = O.lm_head_log_probs(hidden, weight, roll["responses"])
        old = O.perturbed_log_probs(logp, seed=5)
        res = O.fused_loss_reference(hidden[mine], weight, roll["responses"][mine], old[mine], adv[mine],
                                     roll["response_mask"][mine], None)
        # (2b) token-balanced micro-batches: every rank runs the MAXIMUM micro-batch count over the ranks
        # (seqlen_balancing.py:236-239), here rank 1 alone would need fewer
        my_lens = [40, 30, 20, 10, 25, 35] if rank == 0 else [5, 5, 5, 5, 5, 5]
        parts = rearrange_micro_batches(my_lens, 60)
        assert len(parts) == 3 and sorted(i for p in parts for i in p) == list(range(6))
        # (2c) unequal row blocks (speed-aware shards): sizes known to every rank, padded for the collective
        sizes = [3, 5]
        block = torch.arange(sizes[rank] * 2, dtype=torch.float32).view(sizes[rank], 2) + 100 * rank
        got = all_gather_rows(block, sizes=sizes)
        want_rows = torch.cat([torch.arange(n * 2, dtype=torch.float32).view(n, 2) + 100 * r for r, n in enumerate(sizes)])
        assert torch.equal(got, want_rows)
        dw = res["dweight"].clone()
        allreduce_mean_(dw)  # (3) FSDP-style mean over ranks
        if rank == 0:
            want = torch.zeros_like(dw)
            for o in owners:
                want += O.fused_loss_reference(hidden[o], weight, roll["responses"][o], old[o], adv[o],
                                               roll["response_mask"][o], None)["dweight"]
            want /= world
            np.testing.assert_allclose(dw.numpy(), want.numpy(), rtol=1e-5, atol=1e-8)
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_allreduce():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"
