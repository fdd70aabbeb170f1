"""CPU oracle for the GRPO policy-loss hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain PyTorch (fp32, CPU) restatement of the reference's arithmetic for the path named in
BASELINE.json. It exists so the CUDA kernels can be checked against something that does not share any code with them.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it; nothing under ``spatialthinker_b200/`` does (tests/test_no_oracle_in_product.py enforces that).

Parity status: PINNED. The reference (hunarbatra/SpatialThinker) ships no tests or golden vectors for this path, so
the oracle is pinned against outputs of the reference's own functions executed in the build container
(``tests/golden/make_golden.py`` imports ``verl.utils.torch_functional`` and ``verl.trainer.core_algos`` from
``/root/reference`` and writes ``tests/golden/*.npz``), and ``tests/test_oracle.py`` additionally compares the oracle
with the live reference import whenever ``/root/reference`` is present.

Third-party arithmetic on the path that is NOT in the reference tree and is therefore restated from its published
definition: HF ``nn.Linear(hidden, vocab, bias=False)`` (transformers>=4.49, requirements.txt:18) = ``h @ W.T``;
flash-attn's Triton ``cross_entropy_loss`` (flash-attn>=2.4.3, requirements.txt:4) = ``logsumexp(z) - z[label]`` in
fp32; the sign convention is the training branch's ``-CE`` = log p (verl/utils/torch_functional.py:42), not the CPU
fallback's ``+CE`` (:64).

Precision: inputs are bf16-representable values held in fp32; everything after is fp32 ("reference torch GRPO loss
fp32 on CPU", BASELINE.json configs[0]).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

KL_MODES = ("kl", "abs", "mse", "low_var_kl", "chi2")


# ----------------------------------------------------------------------------------------------------------------
# torch_functional surface
# ----------------------------------------------------------------------------------------------------------------
def log_probs_from_logits(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """log p[label] per row, fp32.  verl/utils/torch_functional.py:45-66 with the flash-attn sign (:42)."""
    lead = logits.shape[:-1]
    z = logits.reshape(-1, logits.shape[-1]).float()
    ce = F.cross_entropy(z, labels.reshape(-1), reduction="none")
    return (-ce).view(*lead)


def entropy_from_logits(logits: torch.Tensor) -> torch.Tensor:
    """H = logsumexp(z) - sum softmax(z) * z  (upstream-veRL entropy_from_logits)."""
    z = logits.float()
    return torch.logsumexp(z, dim=-1) - (torch.softmax(z, dim=-1) * z).sum(dim=-1)


def masked_mean(values: torch.Tensor, mask: torch.Tensor, dim: Optional[int] = None, eps: float = 1e-8) -> torch.Tensor:
    """verl/utils/torch_functional.py:69-71."""
    num = (values * mask).sum(dim=dim)
    den = mask.sum(dim=dim) + eps
    return num / den


# ----------------------------------------------------------------------------------------------------------------
# core_algos surface
# ----------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def compute_grpo_outcome_advantage(
    token_level_rewards: torch.Tensor, response_mask: torch.Tensor, index: Sequence, eps: float = 1e-6
) -> Tuple[torch.Tensor, torch.Tensor]:
    """verl/trainer/core_algos.py:137-175.

    Sequence score = sum of token rewards; per uid group the fp32 mean and unbiased fp32 std of the member scores (in
    order of appearance, as the reference's lists are built); a_i = (s_i - mean) / (std + eps); broadcast over the
    response mask. Groups of one sequence are an AssertionError, as in the reference (:167). The same tensor object is
    returned twice (:175).
    """
    seq_score = token_level_rewards.sum(dim=-1)
    members: Dict[object, List[int]] = {}
    for row in range(seq_score.shape[0]):
        members.setdefault(index[row], []).append(row)
    normed = seq_score.clone()
    for uid, rows in members.items():
        assert len(rows) > 1, "GRPO needs rollout.n > 1."
        vals = torch.tensor([seq_score[r] for r in rows])
        mu, sd = torch.mean(vals), torch.std(vals)
        for r in rows:
            normed[r] = (seq_score[r] - mu) / (sd + eps)
    adv = normed.unsqueeze(-1) * response_mask
    return adv, adv


def compute_policy_loss(
    old_log_probs: torch.Tensor,
    log_probs: torch.Tensor,
    advantages: torch.Tensor,
    response_mask: torch.Tensor,
    clip_ratio_low: float,
    clip_ratio_high: float,
    clip_ratio_dual: float,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """verl/trainer/core_algos.py:291-353: (pg_loss, clipfrac_higher, clipfrac_lower, ppo_kl)."""
    x = log_probs - old_log_probs  # "negative_approx_kl" (:331)
    ratio = x.exp()
    lo, hi = np.log(1.0 - clip_ratio_low), np.log(1.0 + clip_ratio_high)  # fp64 scalars, as in :336
    ratio_c = torch.clamp(x, lo, hi).exp()
    l_plain = -advantages * ratio
    l_clip = -advantages * ratio_c
    l_dual = -advantages * clip_ratio_dual
    upper = torch.max(l_plain, l_clip)
    lower = torch.min(upper, l_dual)
    neg_adv = advantages < 0
    per_token = torch.where(neg_adv, lower, upper)
    frac_hi = (l_plain < l_clip).float()
    frac_lo = (upper > l_dual).float() * neg_adv.float()
    return (
        masked_mean(per_token, response_mask),
        masked_mean(frac_hi, response_mask),
        masked_mean(frac_lo, response_mask),
        masked_mean(-x, response_mask),
    )


def compute_kl(log_probs: torch.Tensor, ref_log_probs: torch.Tensor, kl_penalty: str) -> torch.Tensor:
    """verl/trainer/core_algos.py:394-436 (every mode except "full", which needs whole distributions)."""
    lp, ref = log_probs.float(), ref_log_probs.float()
    if kl_penalty == "kl":
        return lp - ref
    if kl_penalty == "abs":
        return (lp - ref).abs()
    if kl_penalty == "mse":
        return 0.5 * (lp - ref).square()
    if kl_penalty == "low_var_kl":
        k = ref - lp
        return torch.clamp(k.exp() - k - 1, min=-10, max=10)
    if kl_penalty == "chi2":
        r = (ref - lp).exp()
        return torch.clamp((r - 1) ** 2, min=0, max=20)
    raise NotImplementedError(f"Unknown KL penalty: {kl_penalty}.")


# ----------------------------------------------------------------------------------------------------------------
# the rest of core_algos / torch_functional / ray_trainer that shares the elementwise skeleton
# ----------------------------------------------------------------------------------------------------------------
def masked_var(values: torch.Tensor, mask: torch.Tensor, unbiased: bool = True) -> torch.Tensor:
    """verl/utils/torch_functional.py:74-89: masked mean of the squared deviations from the masked mean; Bessel's
    correction ``M / (M - 1)`` unless ``M = sum(mask) <= 1`` (the reference warns and keeps the biased value)."""
    dev = values - masked_mean(values, mask)
    var = masked_mean(dev * dev, mask)
    if unbiased:
        m = mask.sum()
        if m > 1:
            var = var * (m / (m - 1))
    return var


def masked_whiten(values: torch.Tensor, mask: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """verl/utils/torch_functional.py:92-95 (every position is transformed, masked or not)."""
    mu = masked_mean(values, mask)
    return (values - mu) * torch.rsqrt(masked_var(values, mask) + eps)


def _group_rows(index: Sequence) -> Dict[object, List[int]]:
    members: Dict[object, List[int]] = {}
    for row, uid in enumerate(index):
        members.setdefault(uid, []).append(row)
    return members


@torch.no_grad()
def compute_rloo_outcome_advantage(token_level_rewards: torch.Tensor, response_mask: torch.Tensor, index: Sequence):
    """verl/trainer/core_algos.py:179-214: score minus the mean of the group's OTHER scores; same object twice."""
    score = token_level_rewards.sum(dim=-1)
    out = score.clone()
    for rows in _group_rows(index).values():
        assert len(rows) > 1, "RLOO needs rollout.n > 1."
        total = torch.sum(torch.tensor([score[r] for r in rows]))
        for r in rows:
            out[r] = score[r] - (total - score[r]) / (len(rows) - 1)
    adv = out.unsqueeze(-1) * response_mask
    return adv, adv


@torch.no_grad()
def compute_remax_outcome_advantage(token_level_rewards: torch.Tensor, reward_baselines: torch.Tensor,
                                    response_mask: torch.Tensor):
    """verl/trainer/core_algos.py:248-273."""
    adv = (token_level_rewards.sum(dim=-1) - reward_baselines).unsqueeze(-1) * response_mask
    return adv, adv


@torch.no_grad()
def compute_reinforce_plus_plus_outcome_advantage(token_level_rewards: torch.Tensor, response_mask: torch.Tensor,
                                                  gamma: float):
    """verl/trainer/core_algos.py:217-245: R_t = r_t + gamma * (R_{t+1} * mask_{t+1}); advantages = whitened R."""
    t_len = token_level_rewards.shape[1]
    returns = torch.empty_like(token_level_rewards)
    tail = torch.zeros_like(token_level_rewards[:, 0])
    for t in range(t_len - 1, -1, -1):
        returns[:, t] = token_level_rewards[:, t] + gamma * tail
        tail = returns[:, t] * response_mask[:, t]
    return masked_whiten(returns, response_mask), returns


@torch.no_grad()
def compute_gae_advantage_return(token_level_rewards: torch.Tensor, values: torch.Tensor, response_mask: torch.Tensor,
                                 gamma: float, lam: float):
    """verl/trainer/core_algos.py:93-133: A_t = (r_t + gamma * v_{t+1} - v_t) + (gamma * lam) * A_{t+1};
    returns = A + v; advantages = whitened A."""
    t_len = token_level_rewards.shape[1]
    adv = torch.empty_like(token_level_rewards)
    nxt_adv = torch.zeros_like(values[:, 0])
    nxt_val = torch.zeros_like(values[:, 0])
    decay = gamma * lam
    for t in range(t_len - 1, -1, -1):
        delta = token_level_rewards[:, t] + gamma * nxt_val - values[:, t]
        nxt_adv = delta + decay * nxt_adv
        adv[:, t] = nxt_adv
        nxt_val = values[:, t]
    return masked_whiten(adv, response_mask), adv + values


def compute_rewards(token_level_scores: torch.Tensor, log_probs: torch.Tensor, ref_log_probs: torch.Tensor,
                    kl_ratio: float) -> torch.Tensor:
    """verl/trainer/core_algos.py:276-283."""
    return token_level_scores - (log_probs - ref_log_probs) * kl_ratio


def compute_value_loss(vpreds: torch.Tensor, returns: torch.Tensor, values: torch.Tensor, action_mask: torch.Tensor,
                       cliprange_value: float):
    """verl/trainer/core_algos.py:356-391: (vf_loss, vf_clipfrac)."""
    clipped = torch.min(torch.max(vpreds, values - cliprange_value), values + cliprange_value)
    err_plain = (vpreds - returns) ** 2
    err_clip = (clipped - returns) ** 2
    loss = 0.5 * masked_mean(torch.max(err_plain, err_clip), action_mask)
    frac = masked_mean((err_plain < err_clip).float(), action_mask)
    return loss, frac


@torch.no_grad()
def kl_penalty_rewards(token_level_scores: torch.Tensor, old_log_probs: Optional[torch.Tensor],
                       ref_log_probs: Optional[torch.Tensor], response_mask: torch.Tensor, kl_coef: float,
                       kl_penalty: str = "kl"):
    """The arithmetic of apply_kl_penalty, verl/trainer/ray_trainer.py:125-145:
    (token_level_rewards, current_kl) with current_kl the batch mean of the per-sequence masked means of kld."""
    if ref_log_probs is not None:
        kld = compute_kl(old_log_probs, ref_log_probs, kl_penalty) * response_mask
    else:
        kld = torch.zeros_like(response_mask, dtype=torch.float32)
    rewards = token_level_scores - kl_coef * kld
    per_seq = masked_mean(kld, response_mask, dim=-1)
    return rewards, float(per_seq.mean(dim=0))


# ----------------------------------------------------------------------------------------------------------------
# lm_head + micro-batch arithmetic (dp_actor.py)
# ----------------------------------------------------------------------------------------------------------------
def lm_head_log_probs(
    hidden: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor, temperature: float = 1.0, want_entropy: bool = False
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """hidden [..., H], weight [V, H] -> (log p[label] [...], entropy [...] or None).

    HF lm_head (``F.linear`` without bias) reached from dp_actor.py:118-125, ``logits.div_(temperature)`` :126,
    log_probs_from_logits :128 - all in fp32 here.
    """
    z = F.linear(hidden.float(), weight.float()) / temperature
    logp = log_probs_from_logits(z, labels)
    ent = entropy_from_logits(z) if want_entropy else None
    return logp, ent


def micro_batch_loss(
    log_probs: torch.Tensor,
    old_log_probs: torch.Tensor,
    advantages: torch.Tensor,
    response_mask: torch.Tensor,
    ref_log_probs: Optional[torch.Tensor],
    *,
    clip_ratio_low: float = 0.2,
    clip_ratio_high: float = 0.3,
    clip_ratio_dual: float = 3.0,
    kl_penalty: str = "low_var_kl",
    kl_coef: float = 1e-2,
    grad_accum: float = 1.0,
    entropy: Optional[torch.Tensor] = None,
    entropy_coef: float = 0.0,
) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """One micro-batch of dp_actor.py:247-286: returns (loss to back-propagate, metrics as 0-d tensors).

    ``actor/pg_loss`` is reported AFTER the KL term has been added (:271, :281), like the reference does.
    ``entropy_coef`` is an extension (0 in the reference, where entropy is only logged, :253).
    """
    entropy_loss = -masked_mean(log_probs, response_mask)
    pg, cf_hi, cf_lo, ppo_kl = compute_policy_loss(
        old_log_probs, log_probs, advantages, response_mask, clip_ratio_low, clip_ratio_high, clip_ratio_dual
    )
    metrics: Dict[str, torch.Tensor] = {}
    total = pg
    if ref_log_probs is not None:
        kl = masked_mean(compute_kl(log_probs, ref_log_probs, kl_penalty), response_mask)
        total = total + kl * kl_coef
        metrics["actor/kl_loss"] = kl.detach()
        metrics["actor/kl_coef"] = torch.tensor(kl_coef)
    if entropy is not None and entropy_coef != 0.0:
        total = total - entropy_coef * masked_mean(entropy, response_mask)
    loss = total / grad_accum
    metrics.update(
        {
            "actor/pg_loss": total.detach(),
            "actor/pg_clipfrac_higher": cf_hi.detach(),
            "actor/pg_clipfrac_lower": cf_lo.detach(),
            "actor/entropy_loss": entropy_loss.detach(),
            "actor/ppo_kl": ppo_kl.detach(),
            "pg_only": pg.detach(),
        }
    )
    return loss, metrics


def fused_loss_reference(
    hidden: torch.Tensor,
    weight: torch.Tensor,
    labels: torch.Tensor,
    old_log_probs: torch.Tensor,
    advantages: torch.Tensor,
    response_mask: torch.Tensor,
    ref_log_probs: Optional[torch.Tensor],
    *,
    temperature: float = 1.0,
    want_grads: bool = True,
    want_entropy: bool = False,
    **loss_kw,
) -> Dict[str, object]:
    """lm_head -> log-probs -> micro_batch_loss -> backward into hidden and weight, all fp32 on the caller's device."""
    h = hidden.detach().float().requires_grad_(want_grads)
    w = weight.detach().float().requires_grad_(want_grads)
    logp, ent = lm_head_log_probs(h, w, labels, temperature, want_entropy or loss_kw.get("entropy_coef", 0.0) != 0.0)
    loss, metrics = micro_batch_loss(logp, old_log_probs, advantages, response_mask, ref_log_probs, entropy=ent, **loss_kw)
    out: Dict[str, object] = {"loss": loss.detach(), "metrics": metrics, "log_probs": logp.detach(),
                              "entropy": None if ent is None else ent.detach()}
    if want_grads:
        loss.backward()
        out["dhidden"], out["dweight"] = h.grad, w.grad
    return out


def update_policy_reference(
    hidden: torch.Tensor,
    weight: torch.Tensor,
    batch: Dict[str, torch.Tensor],
    *,
    global_batch_size_per_device: int,
    micro_batch_size_per_device_for_update: int,
    ppo_epochs: int = 1,
    temperature: float = 1.0,
    **loss_kw,
) -> Dict[str, object]:
    """The mini-/micro-batch loop of dp_actor.py:227-292 over a batch of sequences (hidden is [B, T, H]).

    Returns per-micro-batch metric lists (append_to_dict semantics, py_functional.py:65) and, per mini-batch (one
    optimizer step each), the accumulated gradients of hidden and weight. No optimizer is applied: the loop exists to
    pin the gradient-accumulation arithmetic (GA = global // micro, loss / GA, per-micro-batch normalisation).
    """
    bsz = hidden.shape[0]
    metrics: Dict[str, List[float]] = {}
    steps = []
    ga = global_batch_size_per_device // micro_batch_size_per_device_for_update
    for _ in range(ppo_epochs):
        for s0 in range(0, bsz, global_batch_size_per_device):
            dh_acc = torch.zeros_like(hidden, dtype=torch.float32)
            dw_acc = torch.zeros_like(weight, dtype=torch.float32)
            for m0 in range(s0, min(s0 + global_batch_size_per_device, bsz), micro_batch_size_per_device_for_update):
                sl = slice(m0, min(m0 + micro_batch_size_per_device_for_update, bsz))
                res = fused_loss_reference(
                    hidden[sl], weight, batch["responses"][sl], batch["old_log_probs"][sl], batch["advantages"][sl],
                    batch["response_mask"][sl], batch["ref_log_probs"][sl] if "ref_log_probs" in batch else None,
                    temperature=temperature, grad_accum=float(ga), **loss_kw,
                )
                dh_acc[sl] += res["dhidden"]
                dw_acc += res["dweight"]
                for key, val in res["metrics"].items():
                    if key != "pg_only":
                        metrics.setdefault(key, []).append(float(val))
            steps.append({"dhidden": dh_acc, "dweight": dw_acc})
    return {"metrics": metrics, "steps": steps}


def averaged_clipped_gradient(rank_grads: Sequence[torch.Tensor], max_grad_norm: float,
                              other_sumsq: float = 0.0) -> Dict[str, torch.Tensor]:
    """What every rank's optimizer sees for one replicated bf16 parameter after the reference's optimizer step
    preamble: FSDP averages the ranks' gradients in fp32 (``mp_reduce_dtype``, verl/workers/actor/config.py:58;
    verl/workers/fsdp_workers.py:242-280), ``clip_grad_norm_`` scales by ``min(1, max_norm / (norm + 1e-6))`` with the
    norm over ALL parameters (verl/workers/actor/dp_actor.py:155-167; ``other_sumsq`` = the other parameters' share),
    and the gradient handed to the optimizer has the parameter's dtype.

    The fp32 sum runs in RANK ORDER and is multiplied by the fp32 reciprocal of the world size - the order the peer
    exchange fixes (csrc/peer_kernels.cuh); any other order is an equally valid FSDP result, to fp32 rounding.
    Returns mean (fp32), norm (fp64 scalar), clip (fp32 scalar), grad (bf16).
    """
    world = len(rank_grads)
    mean = rank_grads[0].to(torch.float32).clone()
    for g in rank_grads[1:]:
        mean += g.to(torch.float32)
    mean *= torch.tensor(1.0, dtype=torch.float32) / world
    norm = (mean.double().square().sum() + other_sumsq).sqrt()
    clip = torch.clamp(max_grad_norm / (norm.float() + 1e-6), max=1.0)
    return {"mean": mean, "norm": norm, "clip": clip, "grad": (mean * clip).to(torch.bfloat16)}


# ----------------------------------------------------------------------------------------------------------------
# synthetic inputs - shared by tests, smoke() and bench.py so every arm sees the same tensors
# ----------------------------------------------------------------------------------------------------------------
def synth_head(rows: int, hidden_dim: int, vocab: int, *, seed: int = 0, sigma_w: float = 0.02, device="cpu"):
    """bf16 hidden [rows, H] ~ N(0,1) and weight [V, H] ~ N(0, sigma_w^2)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    hidden = torch.randn(rows, hidden_dim, generator=g).to(torch.bfloat16)
    weight = (sigma_w * torch.randn(vocab, hidden_dim, generator=g)).to(torch.bfloat16)
    return hidden.to(device), weight.to(device)


def synth_rollout(bsz: int, t_len: int, vocab: int, n: int, *, seed: int = 0, ragged: bool = False):
    """labels, response_mask (int64), rewards at the last valid token, permuted uids - all on CPU."""
    assert bsz % n == 0
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    labels = torch.randint(0, vocab, (bsz, t_len), generator=g)
    if ragged:
        u = torch.rand(bsz, generator=g)
        lens = (1 + torch.floor(u * t_len)).clamp(max=t_len).long()
    else:
        lens = torch.full((bsz,), t_len, dtype=torch.long)
    mask = (torch.arange(t_len)[None, :] < lens[:, None]).long()
    # reward = 0.1 f + 0.2 c + 0.5 a + 0.2 s, components Bernoulli / uniform (spatial_sgg.py:653-681 weights)
    comp = torch.rand(bsz, 4, generator=g)
    score = 0.1 * (comp[:, 0] > 0.2).float() + 0.2 * comp[:, 1] + 0.5 * (comp[:, 2] > 0.5).float() + 0.2 * comp[:, 3]
    rewards = torch.zeros(bsz, t_len)
    rewards[torch.arange(bsz), lens - 1] = score
    uid = np.repeat(np.array([f"prompt-{i:05d}" for i in range(bsz // n)], dtype=object), n)
    perm = torch.randperm(bsz, generator=g).numpy()
    return {"responses": labels, "response_mask": mask, "token_level_rewards": rewards, "uid": uid[perm]}


def perturbed_log_probs(logp: torch.Tensor, *, seed: int, jitter: float = 0.1, outlier_frac: float = 0.01):
    """old / ref log-probs: logp + N(0, jitter^2), with a few +-1.5 shifts to reach both clip sides, the dual clip and
    the KL clamp."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = logp.detach().cpu().float() + jitter * torch.randn(logp.shape, generator=g)
    hit = torch.rand(logp.shape, generator=g) < outlier_frac
    sign = torch.where(torch.rand(logp.shape, generator=g) < 0.5, -1.5, 1.5)
    return torch.where(hit, out + sign, out)
