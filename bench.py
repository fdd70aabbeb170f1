"""bench.py - GRPO loss fwd+bwd response-tokens/s on the Qwen2.5-VL-7B head shape (BASELINE.json metric / configs[2]).

    python bench.py --gpus 1 --steps K --warmup W            # this library, one process per GPU (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's own torch functions on the host cores

One "step" = one pass of the hot path over one rollout batch, driven through the drop-in for the reference's actor,
``DataParallelPPOActor.update_policy`` (verl/workers/actor/dp_actor.py:212-292): group-normalised advantages, then for
every micro-batch lm_head -> log-probs -> clipped policy loss + low_var_kl -> dHidden and dW (accumulated in fp32), then
per optimizer step the mean all-reduce of dW over ranks and its norm (dp_actor.py:155-167 minus the optimizer update
itself, which is outside the path), metrics read back to the host. The rollout batch of 512 prompts x n=8 x 1024 response
tokens is sharded by sequence over the ranks (strong scaling: total work fixed); micro-batches are formed by token count
(``use_dynamic_bsz``, 37 888 tokens = two 18 944-row chunks) unless ``--micro-seqs`` asks for the reference's fixed size.

Synthetic inputs follow SURVEY.md section 8(d): ``old_log_probs = logp + 0.1 randn`` with 1 % of the entries shifted by
+-1.5 (both clip sides, the dual clip and the KL clamp are hit), ``ref_log_probs`` likewise, where ``logp`` comes from one
forward pass of the head in setup - so dL/dlogp is non-zero on essentially every token (``--legacy-inputs`` restores
round 1's ``-3 + 0.1 randn``, under which half of the gradient rows are exact zeros).

`value`  : tokens/s with every input already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same pass with the rollout batch in PINNED HOST memory: rewards go host->device, advantages are computed
           and read back, and ``update_policy`` streams every micro-batch's inputs host->device inside the timed region
           (double-buffered on a copy stream); the metrics come back to the host.
`roofline`: the dominant kernel's algorithmic FLOP/s (2*H*V per row) over its CUDA-event duration, live, against the
           measured bf16 tensor peak in MEASURED_PEAKS.json.
`config.records`: the same device-resident pass at the reference's shipped micro-batch of 4 sequences
           (scripts/config.yaml:28) with the deferred dW GEMM, on a few steps.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

CONFIGS = {
    # name: (hidden, vocab, sequences, response_len, group n, description)
    "c3": (3584, 151936, 4096, 1024, 8, "Qwen2.5-VL-7B head bf16, rollout batch 512 x n=8 x 1024 resp tokens, sequence-sharded"),
    "c2": (2048, 151936, 1024, 1024, 8, "Qwen2.5-VL-3B head bf16, rollout batch 128 x n=8 x 1024 resp tokens"),
    "c1": (2048, 151936, 8, 512, 8, "Qwen2.5-VL-3B head, 8 rollouts x 512 resp tokens"),
    "c4": (3584, 151936, 4096, 2048, 8, "Qwen2.5-VL-7B head, ref low_var_kl + entropy output, 2048-token responses"),
    "c5": (3584, 151936, 8192, 4096, 16, "Qwen2.5-VL-7B head, 4096 resp tokens, ragged masks, n=16"),
}
METRIC = "GRPO loss fwd+bwd response-tokens/s (7B head)"
CLIP = dict(clip_ratio_low=0.2, clip_ratio_high=0.3, clip_ratio_dual=3.0)
KL = dict(kl_penalty="low_var_kl", kl_coef=1e-2)
OPT_STEPS = 4  # optimizer steps per rollout step: 4096 sequences / (global batch 128 x n 8), scripts/config.yaml:27
MAX_TOKENS = 37888  # tokens per micro-batch: two 18 944-row chunks of the GEMM pipeline


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"burst": p["bf16_tflops"], "sustained": p["bf16_tflops_sustained"], "hbm": p["hbm_gbs"], "src": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


# ------------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi) for the timed region
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s, p in zip(sm, pw) if p > 300] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
# synthetic workload: the GLOBAL batch structure (lengths, uids, rank assignment) is generated
# identically on every rank, this rank's tensors on its device
# ------------------------------------------------------------------------------------------------------------------
def global_layout(cfg, world, ragged, balance, counts=None, step_times=None):
    """Response lengths and uids of the whole rollout batch in dispatch order: rank r owns rows [offset_r, offset_r + count_r).

    Equal counts (default): token-balanced exactly like the reference's driver does before dispatch (_balance_batch,
    verl/trainer/ray_trainer.py:526-541 -> seqlen_balancing.py:150-181) when the batch is ragged, and additionally per
    optimizer step (balance == 2). ``counts`` / ``step_times`` (speed-aware shards): rank r gets counts[r] sequences and,
    for a ragged batch, tokens in proportion to its measured speed in every mini-batch."""
    from spatialthinker_b200.sharding import balanced_rank_order, speed_weights, weighted_balanced_cells

    _, _, bsz, tlen, n, _ = cfg
    gp = torch.Generator().manual_seed(7)
    uid = torch.arange(bsz // n).repeat_interleave(n)[torch.randperm(bsz, generator=gp)]  # permuted: groups straddle ranks
    if ragged:
        lens = (1 + torch.floor(torch.rand(bsz, generator=gp) * tlen)).clamp(max=tlen).long()
    else:
        lens = torch.full((bsz,), tlen, dtype=torch.long)
    local = bsz // world
    counts = [local] * world if counts is None else list(counts)
    offsets = [sum(counts[:r]) for r in range(world)]
    naive = [int(lens[r * local:(r + 1) * local].sum()) for r in range(world)]
    if ragged and balance and world > 1:
        if step_times is None:
            # balance == 1: the reference's one partition of the whole batch; balance == 2 (default): every optimizer step's
            # mini-batch balanced across the ranks as well (sharding.balanced_rank_order)
            order = torch.tensor(balanced_rank_order(lens.tolist(), world, OPT_STEPS if balance > 1 else 1))
        else:  # speed-aware: every (rank, mini-batch) cell keeps its sequence count and gets tokens in proportion to speed
            sizes = [c // OPT_STEPS for c in counts for _ in range(OPT_STEPS)]
            sw = speed_weights(step_times)
            weights = [sw[r] for r in range(world) for _ in range(OPT_STEPS)]
            cells = weighted_balanced_cells(lens.tolist(), sizes, weights)
            order = torch.tensor([i for cell in cells for i in cell])
        lens, uid = lens[order], uid[order]
    per_rank = [int(lens[offsets[r]:offsets[r] + counts[r]].sum()) for r in range(world)]
    # worst relative deviation, at an optimizer step, of a rank's tokens from its share (equal, or its speed weight)
    share = [1.0] * world if step_times is None else speed_weights(step_times)
    cells = [[float(lens[offsets[r] + m * (counts[r] // OPT_STEPS):offsets[r] + (m + 1) * (counts[r] // OPT_STEPS)].sum()) / share[r]
              for r in range(world)] for m in range(OPT_STEPS)] if all(c >= OPT_STEPS for c in counts) else [[1.0]]
    spread = max((max(c) - min(c)) / max(max(c), 1e-9) for c in cells)
    return lens, uid.numpy(), per_rank, naive, spread, counts, offsets


def make_inputs(st, cfg, rank, world, dev, ragged, legacy, balance, counts=None, step_times=None):
    hdim, vocab, bsz, tlen, n, _ = cfg
    assert bsz % world == 0
    lens_all, uid_all, per_rank, naive, spread, counts, offsets = global_layout(cfg, world, ragged, balance, counts, step_times)
    local = counts[rank]
    lens = lens_all[offsets[rank]:offsets[rank] + local].to(dev)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    gw = torch.Generator(device=dev).manual_seed(99)  # the weight is replicated: same seed on every rank
    weight = torch.empty(vocab, hdim, dtype=torch.bfloat16, device=dev)
    for r0 in range(0, vocab, 16384):
        r1 = min(vocab, r0 + 16384)
        weight[r0:r1] = (0.02 * torch.randn(r1 - r0, hdim, generator=gw, device=dev)).to(torch.bfloat16)
    hidden = torch.empty(local, tlen, hdim, dtype=torch.bfloat16, device=dev)
    blk = max(1, 8192 // tlen)
    for s0 in range(0, local, blk):
        s1 = min(local, s0 + blk)
        hidden[s0:s1] = torch.randn(s1 - s0, tlen, hdim, generator=g, device=dev).to(torch.bfloat16)
    labels = torch.randint(0, vocab, (local, tlen), generator=g, device=dev)
    mask = (torch.arange(tlen, device=dev)[None, :] < lens[:, None]).long()
    # reward = 0.1 f + 0.2 c + 0.5 a + 0.2 s with Bernoulli / uniform components (spatial_sgg.py:653-681), at the last token
    comp = torch.rand(local, 4, generator=g, device=dev)
    score = 0.1 * (comp[:, 0] > 0.2).float() + 0.2 * comp[:, 1] + 0.5 * (comp[:, 2] > 0.5).float() + 0.2 * comp[:, 3]
    rewards = torch.zeros(local, tlen, device=dev)
    rewards[torch.arange(local, device=dev), lens - 1] = score
    if legacy:  # round 1: a plausible level with jitter; every A < 0 token then sits on the clip and has dL/dlogp == 0
        old = -3.0 + 0.1 * torch.randn(local, tlen, generator=g, device=dev)
        ref = -3.0 + 0.1 * torch.randn(local, tlen, generator=g, device=dev)
    else:  # the policy's own log-probs (one forward pass of the head) + jitter + 1 % outliers of +-1.5
        logp = torch.empty(local, tlen, dtype=torch.float32, device=dev)
        blk = max(1, MAX_TOKENS // tlen)
        for s0 in range(0, local, blk):
            logp[s0:s0 + blk], _ = st.fused_lm_head_log_probs(hidden[s0:s0 + blk], weight, labels[s0:s0 + blk], 1.0)

        def perturbed():
            out = logp + 0.1 * torch.randn(local, tlen, generator=g, device=dev)
            hit = torch.rand(local, tlen, generator=g, device=dev) < 0.01
            sign = torch.where(torch.rand(local, tlen, generator=g, device=dev) < 0.5, -1.5, 1.5)
            return torch.where(hit, out + sign, out)

        old, ref = perturbed(), perturbed()
        del logp
    equal = all(c == counts[0] for c in counts)
    return {"weight": weight, "hidden": hidden, "labels": labels, "mask": mask, "rewards": rewards, "uid_np": uid_all,
            "old": old, "ref": ref, "local": local, "lens": lens, "tokens_per_rank": per_rank, "tokens_per_rank_unbalanced": naive,
            "mini_batch_token_spread": spread, "row_begin": offsets[rank], "rank_sizes": None if equal else counts,
            "counts": counts, "nominal_local": bsz // world}


def actor_config(st, x, micro_seqs, want_entropy, entropy_coeff):
    mini = max(x["local"] // OPT_STEPS, 1)
    # speed-aware shards: every rank scales its losses for the NOMINAL mini-batch size, whatever it holds itself
    nominal = 0 if x["rank_sizes"] is None else max(x["nominal_local"] // OPT_STEPS, 1)
    kw = dict(global_batch_size_per_device=mini, loss_scale_batch_size=nominal, use_kl_loss=True, kl_penalty=KL["kl_penalty"],
              kl_coef=KL["kl_coef"], log_true_entropy=want_entropy, entropy_coeff=entropy_coeff, **CLIP)
    if micro_seqs:
        return st.ActorConfig(micro_batch_size_per_device_for_update=micro_seqs, **kw)
    return st.ActorConfig(use_dynamic_bsz=True, max_token_len_per_micro_batch=MAX_TOKENS, **kw)


# ------------------------------------------------------------------------------------------------------------------
# one step through the actor, device-resident inputs
# ------------------------------------------------------------------------------------------------------------------
def run_step_device(st, actor, x, world):
    # advantages: groups straddle ranks, so the per-sequence scores are all-gathered (B floats) and the group statistics
    # run redundantly per rank; this rank's rows are then broadcast over its response mask
    adv, _ = st.core_algos.compute_grpo_outcome_advantage_sharded(x["rewards"], x["mask"], x["uid_np"], x["row_begin"],
                                                                  rank_sizes=x["rank_sizes"])
    data = st.TensorBatch({"hidden_states": x["hidden"], "responses": x["labels"], "response_mask": x["mask"],
                           "old_log_probs": x["old"], "ref_log_probs": x["ref"], "advantages": adv},
                          meta_info={"temperature": 1.0})
    return actor.update_policy(data)


# the round-1 loop (no actor object): kept for A/B against the drop-in (--direct)
def run_step_direct(st, x, plan, dweight, world, want_entropy, entropy_coeff, counts):
    from spatialthinker_b200.dp_actor import grad_sumsq
    from spatialthinker_b200.sharding import allreduce_mean_

    adv, _ = st.core_algos.compute_grpo_outcome_advantage_sharded(x["rewards"], x["mask"], x["uid_np"], x["row_begin"],
                                                                  rank_sizes=x["rank_sizes"])
    metrics, norms = [], []
    for mbs in plan:
        ga = float(len(mbs))
        for sl in mbs:
            res = st.grpo_micro_batch_step(x["hidden"][sl], x["weight"], x["labels"][sl], x["old"][sl], adv[sl], x["ref"][sl],
                                           x["mask"][sl], temperature=1.0, grad_accum=ga, dweight_accum=dweight,
                                           want_entropy=want_entropy, entropy_coeff=entropy_coeff,
                                           valid_rows=None if counts is None else counts[(sl.start, sl.stop)], **CLIP, **KL)
            metrics.append(res["metrics"])
        allreduce_mean_(dweight)
        norms.append(grad_sumsq(dweight, zero_after=True).sqrt())
    return torch.stack(metrics).cpu(), torch.stack(norms).cpu()


def direct_plan(x, micro_seqs):
    local = x["local"]
    mini = max(local // OPT_STEPS, 1)
    plan = [[slice(m0, min(m0 + micro_seqs, s0 + mini)) for m0 in range(s0, s0 + mini, micro_seqs)] for s0 in range(0, local, mini)]
    lens = x["lens"].cpu()
    tlen = x["mask"].shape[1]
    counts = None if bool((lens == tlen).all()) else {(sl.start, sl.stop): int(lens[sl].sum()) for mbs in plan for sl in mbs}
    return plan, counts


# ------------------------------------------------------------------------------------------------------------------
# one step end to end: the rollout batch lives in pinned host memory
# ------------------------------------------------------------------------------------------------------------------
class HostBatch:
    """Pinned host copies of this rank's inputs (what the driver process hands a worker in the reference: a DataProto of
    CPU tensors, moved by ``data.to("cuda")`` in fsdp_workers.py)."""

    def __init__(self, x):
        self.t = {}
        for k in ("hidden", "labels", "old", "ref", "mask", "rewards"):
            buf = torch.empty(x[k].shape, dtype=x[k].dtype, pin_memory=True)
            buf.copy_(x[k])
            self.t[k] = buf
        self.adv = torch.empty(x["rewards"].shape, dtype=torch.float32, pin_memory=True)
        self.rewards_dev = torch.empty_like(x["rewards"])
        self.mask_dev = torch.empty_like(x["mask"])


def run_step_e2e(st, actor, x, host, world):
    # rewards + mask host -> device, advantages on the device, back to the host batch (the reference computes them on the
    # driver and ships them with the batch: ray_trainer.py:148-175)
    host.rewards_dev.copy_(host.t["rewards"], non_blocking=True)
    host.mask_dev.copy_(host.t["mask"], non_blocking=True)
    adv, _ = st.core_algos.compute_grpo_outcome_advantage_sharded(host.rewards_dev, host.mask_dev, x["uid_np"], x["row_begin"],
                                                                  rank_sizes=x["rank_sizes"])
    host.adv.copy_(adv, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    data = st.TensorBatch({"hidden_states": host.t["hidden"], "responses": host.t["labels"], "response_mask": host.t["mask"],
                           "old_log_probs": host.t["old"], "ref_log_probs": host.t["ref"], "advantages": host.adv},
                          meta_info={"temperature": 1.0})
    return actor.update_policy(data)  # streams every micro-batch host -> device; metrics come back as python floats


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own torch functions (baseline/_ref, the unmodified package installed by build()) on the
# host cores; the oracle port when that install is absent
# ------------------------------------------------------------------------------------------------------------------
def load_reference():
    """(VF, core_algos) of the UNMODIFIED reference from baseline/_ref, or None."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "verl")):
        return None
    sys.dont_write_bytecode = True
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    try:
        import verl.trainer.core_algos as ca
        import verl.utils.torch_functional as VF
    except Exception:
        return None
    VF.FLAH_ATTN_CROSS_ENTROPY_LOSS_AVAILABLE = False  # the Triton kernel needs CUDA tensors; this arm runs on the host
    return VF, ca


def cpu_pass(cfg, tokens, seed=0, ref=None):
    """One bounded sample of the path on the host: advantages -> lm_head -> log-probs -> loss -> backward, fp32."""
    import torch.nn.functional as F

    from oracle import grpo_oracle as O  # synthetic inputs (and the arithmetic, when the reference is not installed)

    hdim, vocab, _, _, n, _ = cfg
    seqs = n
    tlen = max(tokens // seqs, 1)
    hid, w = O.synth_head(seqs * tlen, hdim, vocab, seed=seed)
    hid = hid.view(seqs, tlen, hdim)
    roll = O.synth_rollout(seqs, tlen, vocab, n, seed=seed)
    with torch.no_grad():  # old / ref around the policy's own log-probs, as in the GPU arm (not timed)
        lp0, _ = O.lm_head_log_probs(hid, w, roll["responses"])
    old, ref_lp = O.perturbed_log_probs(lp0, seed=seed + 1), O.perturbed_log_probs(lp0, seed=seed + 2)
    mask = roll["response_mask"]
    t0 = time.perf_counter()
    if ref is None:
        adv, _ = O.compute_grpo_outcome_advantage(roll["token_level_rewards"].clone(), mask, roll["uid"])
        O.fused_loss_reference(hid, w, roll["responses"], old, adv, mask, ref_lp, kl_penalty="low_var_kl", kl_coef=1e-2,
                               grad_accum=1.0, **CLIP)
    else:
        VF, ca = ref
        adv, _ = ca.compute_grpo_outcome_advantage(roll["token_level_rewards"].clone(), mask, roll["uid"])
        h = hid.float().requires_grad_(True)
        wf = w.float().requires_grad_(True)
        z = F.linear(h, wf) / 1.0  # HF lm_head (nn.Linear, third party) + logits.div_(temperature), dp_actor.py:125-126
        logp = -VF.log_probs_from_logits(z, roll["responses"])  # the CPU branch returns +CE (torch_functional.py:64)
        pg, _, _, _ = ca.compute_policy_loss(old, logp, adv, mask, CLIP["clip_ratio_low"], CLIP["clip_ratio_high"],
                                             CLIP["clip_ratio_dual"])
        kl = VF.masked_mean(ca.compute_kl(logp, ref_lp, "low_var_kl"), mask)
        ((pg + 1e-2 * kl) / 1.0).backward()  # dp_actor.py:271-278
    return seqs * tlen, time.perf_counter() - t0


def cpu_baseline(cfg, budget_s=15.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = load_reference()
    tokens = 512
    cpu_pass(cfg, tokens, ref=ref)  # also warms the thread pool
    total_tok, total_t, passes = 0, 0.0, 0
    while total_t < budget_s and passes < 8:
        n_tok, dt = cpu_pass(cfg, tokens, seed=passes + 1, ref=ref)
        total_tok += n_tok
        total_t += dt
        passes += 1
    what = ("verl.utils.torch_functional + verl.trainer.core_algos of the unmodified reference (baseline/_ref) around torch F.linear"
            if ref is not None else "oracle/grpo_oracle.py (reference not installed)")
    return {"value": total_tok / total_t, "unit": "tokens/s", "cores": torch.get_num_threads(),
            "kind": "reference" if ref is not None else "port",
            "sample": f"{passes} passes x {tokens} response tokens ({cfg[4]} sequences) of the same head shape, fp32 torch on "
                      f"{cores} host threads; {what}"}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = load_reference()
    tokens = 512
    for _ in range(max(args.warmup, 1)):
        cpu_pass(cfg, tokens, ref=ref)
    tot_tok, tot_t = 0, 0.0
    for i in range(args.steps):
        n_tok, dt = cpu_pass(cfg, tokens, seed=i, ref=ref)
        tot_tok += n_tok
        tot_t += dt
    val = tot_tok / tot_t
    kind = "reference" if ref is not None else "port"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {cfg[5]}; each step = a bounded sample of {tokens} response tokens on the host"},
        "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"{args.steps} steps x {tokens} tokens; " +
                                   ("the unmodified reference's torch_functional / core_algos (baseline/_ref) around torch F.linear"
                                    if ref is not None else "oracle port of the reference torch path")},
        "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--micro-seqs", type=int, default=0,
                    help="fixed sequences per micro-batch as in the reference (0: token-balanced micro-batches of 37888 tokens)")
    ap.add_argument("--sequences", type=int, default=0, help="override the rollout batch size (debug)")
    ap.add_argument("--entropy-coeff", type=float, default=0.0,
                    help="loss -= coeff * masked_mean(entropy): adds the per-element stash -> dlogits pass (not in the reference)")
    ap.add_argument("--no-defer-dw", action="store_true", help="actor without the deferred dW GEMM for small micro-batches")
    ap.add_argument("--direct", action="store_true", help="round-1 loop over grpo_micro_batch_step instead of the actor (A/B)")
    ap.add_argument("--legacy-inputs", action="store_true", help="round-1 old/ref log-probs (-3 + 0.1 randn)")
    ap.add_argument("--balance", type=int, default=2, choices=[0, 1, 2],
                    help="ragged config: 0 = contiguous rank shards, 1 = the reference's token-balanced shards (_balance_batch), "
                         "2 = every optimizer step's mini-batch balanced across the ranks as well")
    ap.add_argument("--no-records", action="store_true", help="skip the extra record at 4-sequence micro-batches")
    ap.add_argument("--no-speed-aware", action="store_true",
                    help="N > 1: keep equal sequence counts per rank (default: re-deal sequences by measured per-rank speed)")
    ap.add_argument("--peer", default="auto", choices=["auto", "on", "off"],
                    help="N > 1: dW exchange over peer-mapped memory (this library's kernels over NVLink; auto = on when the "
                         "GPUs can map each other's memory) or NCCL's fp32 all-reduce (off)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    cfg = list(CONFIGS[args.config])
    if args.sequences:
        cfg[2] = args.sequences
    cfg = tuple(cfg)
    if args.impl == "reference":
        run_reference(args, cfg)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL printf()s its banner ("NCCL version ...") when the first communicator
        # is created, so file descriptor 1 points at stderr until that has happened
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    import spatialthinker_b200 as st
    from spatialthinker_b200 import _lib

    lib = st.load_library()
    hdim, vocab, bsz, tlen, n, desc = cfg
    ragged = args.config == "c5"
    want_entropy = args.config == "c4"
    x = make_inputs(st, cfg, rank, world, dev, ragged, args.legacy_inputs, args.balance)

    def count_tokens(xx):
        local = int(xx["mask"].sum().item())
        tok = torch.tensor([local], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tok)
        return local, float(tok.item())

    tokens_local, tokens_total = count_tokens(x)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def build_actor(xx, micro_seqs=args.micro_seqs, defer=not args.no_defer_dw):
        return st.DataParallelPPOActor(actor_config(st, xx, micro_seqs, want_entropy, args.entropy_coeff), xx["weight"],
                                       defer_dw=defer, peer_exchange={"auto": None, "on": True, "off": False}[args.peer])

    actor = build_actor(x)
    # ---- speed-aware shards (N > 1): the chips of one box differ by several percent in sustained throughput under the
    # power cap and meet at every optimizer step; the first warm-up steps on EQUAL shards measure every rank's own kernel
    # time per step, then sequences are re-dealt in proportion to speed (sharding.speed_weighted_counts) and one more
    # warm-up step runs on the new shards. All of it before the timed region.
    speed_aware = None
    warm_done = 0
    if world > 1 and not args.no_speed_aware and not args.direct and args.warmup >= 3:
        run_step_device(st, actor, x, world)
        lib.grpo_profile_enable(1)
        lib.grpo_profile_read(None, None, 1)
        for _ in range(args.warmup - 2):
            run_step_device(st, actor, x, world)
        torch.cuda.synchronize()
        cal_ms = (ctypes.c_double * _lib.NUM_PHASES)()
        lib.grpo_profile_read(cal_ms, None, 1)
        lib.grpo_profile_enable(0)
        mine = torch.tensor([sum(cal_ms) / (args.warmup - 2)], dtype=torch.float64, device=dev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        times = [float(t.item()) for t in every]
        from spatialthinker_b200.sharding import speed_weighted_counts

        # dense responses: the only way to move work is to move sequences; ragged ones: every rank keeps its sequence count
        # (the reference's equal-size dispatch) and the faster ranks get the longer sequences
        counts = [bsz // world] * world if ragged else speed_weighted_counts(bsz, times, OPT_STEPS)
        speed_aware = {"own_kernel_ms_per_step_on_equal_shards": [round(t, 1) for t in times], "sequences_per_rank": counts,
                       "how": "tokens per rank and mini-batch in proportion to speed, equal sequence counts" if ragged
                              else "sequences per rank in proportion to speed"}
        warm_done = args.warmup - 1
        if ragged or any(c != counts[0] for c in counts):
            actor.release_workspaces()
            del actor, x
            torch.cuda.empty_cache()
            x = make_inputs(st, cfg, rank, world, dev, ragged, args.legacy_inputs, args.balance, counts, times)
            tokens_local, tokens_total = count_tokens(x)
            actor = build_actor(x)
    if args.direct:
        micro = args.micro_seqs or max(1, MAX_TOKENS // tlen)
        plan, counts = direct_plan(x, micro)
        dweight = torch.zeros(vocab, hdim, dtype=torch.float32, device=dev)
        dev_step = lambda: run_step_direct(st, x, plan, dweight, world, want_entropy, args.entropy_coeff, counts)  # noqa: E731
    else:
        seen = {}
        dev_step = lambda: seen.update(m=run_step_device(st, actor, x, world))  # noqa: E731
    for _ in range(args.warmup - warm_done):
        dev_step()
    actor.time_collectives = world > 1
    actor.collective_events = []
    sampler = ClockSampler(local_rank)
    launches0 = lib.grpo_launch_count()
    lib.grpo_profile_enable(1)
    lib.grpo_profile_read(None, None, 1)
    if rank == 0:
        sampler.start()
    ms_total = timed(dev_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.grpo_launch_count() - launches0
    ph_ms = (ctypes.c_double * _lib.NUM_PHASES)()
    ph_cnt = (ctypes.c_longlong * _lib.NUM_PHASES)()
    lib.grpo_profile_read(ph_ms, ph_cnt, 1)
    lib.grpo_profile_enable(0)
    ms_per_step = ms_total / args.steps
    value = tokens_total / (ms_per_step * 1e-3)
    # per rank: time inside the dW all-reduces per step (wire + waiting for the slowest rank) and in this rank's own GEMMs
    actor.time_collectives = False
    ar_ms = sum(a.elapsed_time(b) for a, b in actor.collective_events) / args.steps
    per_rank = torch.tensor([ar_ms, sum(ph_ms[i] for i in range(_lib.NUM_PHASES)) / args.steps], dtype=torch.float64, device=dev)
    gathered = [torch.zeros_like(per_rank) for _ in range(world)]
    if world > 1:
        dist.all_gather(gathered, per_rank)
    else:
        gathered = [per_rank]
    by_rank = {"allreduce_ms_per_step": [round(float(g[0]), 2) for g in gathered],
               "kernel_phase_ms_per_step": [round(float(g[1]), 1) for g in gathered]}
    n_micro = None if args.direct else len(seen["m"]["actor/pg_loss"])
    if world == 1:
        dw_exchange = "none (one GPU)"
    elif not args.direct and actor._peer:
        dw_exchange = ("peer-mapped memory, own kernels over NVLink: fp32 reduce-scatter + norm, clip + bf16 all-gather + zero "
                       "(spatialthinker_b200/peer.py)")
    else:
        dw_exchange = "NCCL fp32 all-reduce (AVG), then norm + zero"

    # ---- the reference's shipped micro-batch size (4 sequences) with the deferred dW GEMM, a few steps
    records = []
    if not args.no_records and not args.direct and not args.micro_seqs and (x["nominal_local"] // OPT_STEPS) % 4 == 0:
        actor4 = build_actor(x, micro_seqs=4, defer=True)
        step4 = lambda: run_step_device(st, actor4, x, world)  # noqa: E731
        step4()
        k = max(2, args.steps // 10)
        ms4 = timed(step4, k) / k
        records.append({"name": "reference micro-batch size", "micro_batch_sequences": 4, "defer_dw": True, "via": "DataParallelPPOActor.update_policy",
                        "steps": k, "warmup": 1, "ms_per_step": ms4, "value": tokens_total / (ms4 * 1e-3), "unit": "tokens/s",
                        "vs_default": ms_per_step / ms4})
        actor4.release_workspaces()
        del actor4

    # ---- end to end from pinned host memory
    e2e = None
    if not args.no_e2e and not args.direct:
        host = HostBatch(x)
        e2e_step = lambda: run_step_e2e(st, actor, x, host, world)  # noqa: E731
        for _ in range(min(args.warmup, 1)):
            e2e_step()
        h2d0 = actor._stager.h2d_bytes
        ms_e2e = timed(e2e_step, args.steps) / args.steps
        h2d = (actor._stager.h2d_bytes - h2d0) // args.steps + host.t["rewards"].numel() * 4 + host.t["mask"].numel() * 8
        d2h = host.adv.numel() * 4 + (n_micro or 0) * _lib.NUM_METRICS * 4 + OPT_STEPS * 4
        e2e = {"value": tokens_total / (ms_e2e * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": int(h2d * world),
               "d2h_bytes_per_step": int(d2h * world), "ms_per_step": ms_e2e,
               "via": "DataParallelPPOActor.update_policy on a pinned-host batch (advantage kernels and their D2H included)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    kernels = []
    for i, name in enumerate(_lib.PHASE_NAMES):
        if ph_cnt[i]:
            kernels.append({"name": name, "launches": int(ph_cnt[i]), "avg_ms": ph_ms[i] / ph_cnt[i], "total_ms": ph_ms[i]})
    gemms = [k for k in kernels if k["name"].endswith("_gemm")]
    dom = max(gemms, key=lambda k: k["total_ms"])
    rows_per_launch = tokens_local * args.steps / dom["launches"]  # every GEMM launch covers one chunk of (valid) rows (rank 0's)
    achieved = 2.0 * hdim * vocab * rows_per_launch / (dom["avg_ms"] * 1e-3) / 1e12
    peak = peaks["sustained"]  # kernels are timed inside a seconds-long step under the power cap
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and (hdim, vocab) == (3584, 151936):  # the ncu capture was taken at this head shape
        t = json.load(open(tpath)).get(dom["name"])
        if t:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    gemm_ms = sum(k["total_ms"] for k in gemms) / args.steps
    roofline = {"bound": "tensor", "kernel": dom["name"], "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "peak_kind": f"{peaks['src']} sustained bf16 (burst {peaks['burst']})",
                "traffic": traffic, "algorithmic_flops_per_launch": 2.0 * hdim * vocab * rows_per_launch, "kernels": kernels,
                "ms_per_step_outside_phase_timers": ms_per_step - sum(k["total_ms"] for k in kernels) / args.steps,
                "gemm_ms_per_step": gemm_ms,
                "whole_step_tflops_algorithmic": 6.0 * hdim * vocab * tokens_total / (ms_per_step * 1e-3) / 1e12 / world}
    micro_desc = (f"{args.micro_seqs} sequences" if args.micro_seqs else
                  f"token-balanced, <= {MAX_TOKENS} tokens ({n_micro} micro-batches per step)" if not args.direct else
                  f"{max(1, MAX_TOKENS // tlen)} sequences (direct loop)")
    out = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"{args.config}: {desc}", "hidden": hdim, "vocab": vocab, "sequences": bsz, "response_len": tlen,
                   "group_n": n, "via": "grpo_micro_batch_step loop" if args.direct else "DataParallelPPOActor.update_policy",
                   "micro_batch": micro_desc, "optimizer_steps_per_step": OPT_STEPS, "defer_dw": not args.no_defer_dw,
                   "inputs": "round-1 (-3 + 0.1 randn)" if args.legacy_inputs else "SURVEY 8(d): old/ref = logp + 0.1 randn, 1% +-1.5 outliers",
                   "loss": "GRPO clip .2/.3/3.0 + low_var_kl 1e-2" + (f" - {args.entropy_coeff} * entropy" if args.entropy_coeff else ""),
                   "l2": "inputs (>= 30 GB) far exceed the 126 MB L2",
                   "parallelism": f"dp{world} by sequence" + ((", token-balanced rank shards (Karmarkar-Karp" + (", per optimizer step" if args.balance > 1 else "") + ")") if ragged and args.balance else "") + ", dW averaged over ranks once per optimizer step",
                   "dw_exchange": dw_exchange,
                   "tokens_per_rank": x["tokens_per_rank"], "tokens_per_rank_unbalanced": x["tokens_per_rank_unbalanced"],
                   "mini_batch_token_spread_across_ranks": x["mini_batch_token_spread"],
                   "speed_aware_shards": speed_aware, "records": records, "by_rank": by_rank},
        "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e,
    }
    if not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(cfg)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
