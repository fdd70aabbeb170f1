"""Drop-in for the GRPO part of ``verl/trainer/core_algos.py``.

``compute_grpo_outcome_advantage`` (reference :137-175), ``compute_policy_loss`` (:291-353) and ``compute_kl`` (:394-436,
also exported as ``kl_penalty``, the upstream-veRL name) keep the reference's signatures, return conventions and
error behaviour; the arithmetic runs in ``csrc/advantage_kernels.cuh`` / ``csrc/loss_kernels.cuh`` through the C ABI.

The rest of the reference file rides on the same elementwise / warp-per-sequence skeleton (SURVEY.md §8 f-4,
``csrc/estimator_kernels.cuh``): ``compute_gae_advantage_return`` (:93-133), ``compute_rloo_outcome_advantage``
(:179-214), ``compute_reinforce_plus_plus_outcome_advantage`` (:217-245), ``compute_remax_outcome_advantage``
(:248-273), ``compute_rewards`` (:276-283), ``compute_value_loss`` (:356-391) and the host-only KL controllers
(:36-89).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._util import f32c, mask_arg, require_cuda


# ----------------------------------------------------------------------------------------------------------------
# advantages
# ----------------------------------------------------------------------------------------------------------------
def group_csr(index: Sequence) -> Tuple[np.ndarray, np.ndarray]:
    """uid per row (any hashables / numpy object array of strings, arbitrary order) -> (order, offsets) int32 arrays:
    ``order[offsets[g]:offsets[g+1]]`` are the rows of group ``g``, in order of appearance.

    Raises AssertionError for a group of one sequence, exactly like the reference (core_algos.py:167).
    """
    if isinstance(index, torch.Tensor):
        index = index.detach().cpu().numpy()
    idx = np.asarray(index)
    if idx.ndim != 1:
        raise ValueError("index must be one-dimensional")
    if idx.dtype == object:
        idx = idx.astype(str)
    _, inverse, counts = np.unique(idx, return_inverse=True, return_counts=True)
    assert counts.size == 0 or counts.min() > 1, "GRPO needs rollout.n > 1."
    order = np.argsort(inverse, kind="stable").astype(np.int32)
    offsets = np.zeros(counts.size + 1, dtype=np.int32)
    np.cumsum(counts, out=offsets[1:])
    return order, offsets


@torch.no_grad()
def compute_grpo_outcome_advantage(
    token_level_rewards: torch.Tensor, response_mask: torch.Tensor, index: Sequence, eps: float = 1e-6
) -> Tuple[torch.Tensor, torch.Tensor]:
    """GRPO outcome advantage. Reference: verl/trainer/core_algos.py:137-175 (caller ray_trainer.py:148-175).

    Args:
        token_level_rewards: (bs, response_length) float
        response_mask: (bs, response_length), any of int64 / float32 / bool
        index: uid per sequence (numpy object array of strings in the reference), groups in arbitrary row order

    Returns:
        (advantages, returns), both (bs, response_length) float32 and - as in the reference - the same tensor object.
    """
    dev = require_cuda(token_level_rewards, response_mask)
    if token_level_rewards.dim() != 2 or response_mask.shape != token_level_rewards.shape:
        raise ValueError("token_level_rewards and response_mask must both be (bs, response_length)")
    bsz, t_len = token_level_rewards.shape
    if len(index) != bsz:
        raise ValueError(f"index has {len(index)} entries for a batch of {bsz}")
    order, offsets = group_csr(index)
    lib = _lib.load()
    rewards = f32c(token_level_rewards)
    mask, code = mask_arg(response_mask)
    order_d = torch.from_numpy(order).to(dev, non_blocking=True)
    offsets_d = torch.from_numpy(offsets).to(dev, non_blocking=True)
    adv = torch.empty(bsz, t_len, dtype=torch.float32, device=dev)
    seq = torch.empty(2 * max(bsz, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            lib.grpo_advantage(rewards.data_ptr(), mask.data_ptr(), code, order_d.data_ptr(), offsets_d.data_ptr(), bsz,
                               t_len, offsets.size - 1, float(eps), adv.data_ptr(), seq.data_ptr(),
                               _lib.stream_ptr(dev)),
            "grpo_advantage",
        )
    return adv, adv


@torch.no_grad()
def compute_grpo_outcome_advantage_sharded(
    token_level_rewards: torch.Tensor,
    response_mask: torch.Tensor,
    index_all: Sequence,
    row_begin: int,
    eps: float = 1e-6,
    group=None,
    rank_sizes: Optional[Sequence[int]] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """The same advantages when the batch is sharded by sequence over data-parallel ranks.

    ``token_level_rewards`` / ``response_mask`` are this rank's rows ``[row_begin, row_begin + bs_local)`` of the global
    batch (rank order = row order; equal row counts per rank unless ``rank_sizes`` lists them - speed-aware shards);
    ``index_all`` is the uid of EVERY sequence. Scores are all-gathered (``bs_all`` floats over NCCL), the group
    statistics run redundantly on every rank.
    """
    from .sharding import all_gather_rows

    dev = require_cuda(token_level_rewards, response_mask)
    bsz, t_len = token_level_rewards.shape
    order, offsets = group_csr(index_all)
    lib = _lib.load()
    rewards = f32c(token_level_rewards)
    mask, code = mask_arg(response_mask)
    scores = torch.empty(bsz, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.grpo_sequence_scores(rewards.data_ptr(), bsz, t_len, scores.data_ptr(), _lib.stream_ptr(dev)),
                   "grpo_sequence_scores")
    scores_all = all_gather_rows(scores, group, rank_sizes)
    bsz_all = scores_all.shape[0]
    if bsz_all != len(index_all):
        raise ValueError(f"index_all has {len(index_all)} entries but the gathered batch has {bsz_all} sequences")
    order_d = torch.from_numpy(order).to(dev, non_blocking=True)
    offsets_d = torch.from_numpy(offsets).to(dev, non_blocking=True)
    adv = torch.empty(bsz, t_len, dtype=torch.float32, device=dev)
    seq = torch.empty(max(bsz_all, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            lib.grpo_advantage_from_scores(scores_all.data_ptr(), order_d.data_ptr(), offsets_d.data_ptr(), bsz_all,
                                           offsets.size - 1, float(eps), int(row_begin), mask.data_ptr(), code, bsz,
                                           t_len, adv.data_ptr(), seq.data_ptr(), _lib.stream_ptr(dev)),
            "grpo_advantage_from_scores",
        )
    return adv, adv


# ----------------------------------------------------------------------------------------------------------------
# policy loss
# ----------------------------------------------------------------------------------------------------------------
class _PolicyLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, old_log_probs, log_probs, advantages, response_mask, clip_low, clip_high, clip_dual):
        dev = require_cuda(old_log_probs, log_probs, advantages, response_mask)
        lib = _lib.load()
        shape = log_probs.shape
        if old_log_probs.shape != shape or response_mask.shape != shape:
            raise ValueError("old_log_probs, log_probs and response_mask must have the same shape")
        lp, old = f32c(log_probs), f32c(old_log_probs)
        adv = f32c(advantages.expand(shape) if advantages.shape != shape else advantages)
        mask, code = mask_arg(response_mask)
        n = lp.numel()
        dlogp = torch.empty(n, dtype=torch.float32, device=dev)
        metrics = torch.empty(_lib.NUM_METRICS, dtype=torch.float32, device=dev)
        acc = torch.empty(16, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(
                lib.grpo_policy_loss_fwd_bwd(lp.data_ptr(), old.data_ptr(), adv.data_ptr(), None, mask.data_ptr(), code,
                                             n, float(clip_low), float(clip_high), float(clip_dual), -1, 0.0, 1.0,
                                             dlogp.data_ptr(), metrics.data_ptr(), acc.data_ptr(),
                                             _lib.stream_ptr(dev)),
                "grpo_policy_loss_fwd_bwd",
            )
        ctx.save_for_backward(dlogp, mask, metrics)
        ctx.shape, ctx.dtype = shape, log_probs.dtype
        out = (metrics[_lib.MET_PG_LOSS], metrics[_lib.MET_CLIPFRAC_HI], metrics[_lib.MET_CLIPFRAC_LO],
               metrics[_lib.MET_PPO_KL])
        ctx.mark_non_differentiable(out[1], out[2])
        return out

    @staticmethod
    def backward(ctx, g_pg, g_hi, g_lo, g_kl):
        dlogp, mask, metrics = ctx.saved_tensors
        grad = None
        if g_pg is not None:
            grad = g_pg * dlogp.view(ctx.shape)
        if g_kl is not None:  # ppo_kl = masked_mean(-(logp - old)) is differentiable in the reference too
            denom = metrics[_lib.MET_MASK_SUM] + 1e-8
            extra = -g_kl * mask.view(ctx.shape).float() / denom
            grad = extra if grad is None else grad + extra
        if grad is not None:
            grad = grad.to(ctx.dtype)
        return None, grad, None, None, None, None, None


def compute_policy_loss(
    old_log_probs: torch.Tensor,
    log_probs: torch.Tensor,
    advantages: torch.Tensor,
    response_mask: torch.Tensor,
    clip_ratio_low: float,
    clip_ratio_high: float,
    clip_ratio_dual: float,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Clipped / dual-clipped policy-gradient loss. Reference: verl/trainer/core_algos.py:291-353.

    Returns 0-d tensors (pg_loss, pg_clipfrac_higher, pg_clipfrac_lower, ppo_kl); ``pg_loss`` and ``ppo_kl`` are
    differentiable with respect to ``log_probs``.
    """
    return _PolicyLoss.apply(old_log_probs, log_probs, advantages, response_mask, clip_ratio_low, clip_ratio_high,
                             clip_ratio_dual)


# ----------------------------------------------------------------------------------------------------------------
# KL estimators
# ----------------------------------------------------------------------------------------------------------------
class _ComputeKL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs, ref_log_probs, mode: int):
        dev = require_cuda(log_probs, ref_log_probs)
        lib = _lib.load()
        if ref_log_probs.shape != log_probs.shape:
            raise ValueError("log_probs and ref_log_probs must have the same shape")
        lp, ref = f32c(log_probs), f32c(ref_log_probs)
        out = torch.empty_like(lp)
        need_grad = log_probs.requires_grad or ref_log_probs.requires_grad
        dout = torch.empty_like(lp) if need_grad else None
        with torch.cuda.device(dev):
            _lib.check(lib.grpo_compute_kl(lp.data_ptr(), ref.data_ptr(), lp.numel(), mode, out.data_ptr(),
                                           _lib.ptr(dout), _lib.stream_ptr(dev)), "grpo_compute_kl")
        if need_grad:
            ctx.save_for_backward(dout)
        ctx.dtypes = (log_probs.dtype, ref_log_probs.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        (dout,) = ctx.saved_tensors
        gl = g * dout  # every estimator is a function of (log_probs - ref_log_probs): d/dref = -d/dlogp
        return gl.to(ctx.dtypes[0]), (-gl).to(ctx.dtypes[1]), None


def compute_kl(log_probs: torch.Tensor, ref_log_probs: torch.Tensor, kl_penalty: str) -> torch.Tensor:
    """Per-token KL estimator, fp32. Reference: verl/trainer/core_algos.py:394-436.

    Modes: "kl", "abs", "mse", "low_var_kl" (shipped default, scripts/config.yaml:22), "chi2". "full" needs whole
    distributions rather than token log-probs and is not on the GRPO path; like any unknown name it raises
    NotImplementedError with the reference's message.
    """
    mode = _lib.KL_MODES.get(kl_penalty) if isinstance(kl_penalty, str) else None
    if mode is None or mode < 0:
        raise NotImplementedError(f"Unknown KL penalty: {kl_penalty}.")
    return _ComputeKL.apply(log_probs, ref_log_probs, mode)


def kl_penalty(logprob: torch.Tensor, ref_logprob: torch.Tensor, kl_penalty: str) -> torch.Tensor:  # noqa: F811
    """Upstream-veRL name for :func:`compute_kl` (BASELINE.json north_star spelling)."""
    return compute_kl(logprob, ref_logprob, kl_penalty)


# ----------------------------------------------------------------------------------------------------------------
# KL controllers (host-only bookkeeping, verl/trainer/core_algos.py:36-89)
# ----------------------------------------------------------------------------------------------------------------
class KLController(ABC):
    kl_coef: float

    @abstractmethod
    def update(self, current_kl: float, n_steps: int) -> None:
        ...


class AdaptiveKLController(KLController):
    """Proportional controller of https://arxiv.org/abs/1909.08593 (reference :46-60): the coefficient moves by at
    most 20 % of ``n_steps / horizon`` per update towards ``target_kl``."""

    def __init__(self, init_kl_coef: float, target_kl: float, horizon: float):
        self.kl_coef = init_kl_coef
        self.target = target_kl
        self.horizon = horizon

    def update(self, current_kl: float, n_steps: int) -> None:
        error = float(np.clip(current_kl / self.target - 1, -0.2, 0.2))
        self.kl_coef *= 1 + error * n_steps / self.horizon


class FixedKLController(KLController):
    """Reference :63-72."""

    def __init__(self, init_kl_coef: float):
        self.kl_coef = init_kl_coef

    def update(self, current_kl: float, n_steps: int) -> None:
        pass


def get_kl_controller(algorithm_config) -> KLController:
    """Reference :75-89: ``kl_type`` "fixed" | "adaptive" (needs ``kl_horizon > 0``), anything else is a ValueError."""
    if algorithm_config.kl_type == "fixed":
        return FixedKLController(init_kl_coef=algorithm_config.kl_coef)
    if algorithm_config.kl_type == "adaptive":
        assert algorithm_config.kl_horizon > 0, f"horizon must be larger than 0. Got {algorithm_config.kl_horizon}."
        return AdaptiveKLController(init_kl_coef=algorithm_config.kl_coef, target_kl=algorithm_config.kl_target,
                                    horizon=algorithm_config.kl_horizon)
    raise ValueError(f"Unknown kl type: {algorithm_config.kl_type}.")


# ----------------------------------------------------------------------------------------------------------------
# the other advantage estimators
# ----------------------------------------------------------------------------------------------------------------
def _seq_inputs(token_level_rewards: torch.Tensor, response_mask: torch.Tensor):
    dev = require_cuda(token_level_rewards, response_mask)
    if token_level_rewards.dim() != 2 or response_mask.shape != token_level_rewards.shape:
        raise ValueError("token_level_rewards and response_mask must both be (bs, response_length)")
    mask, code = mask_arg(response_mask)
    return dev, f32c(token_level_rewards), mask, code


def group_csr_any(index: Sequence, what: str) -> Tuple[np.ndarray, np.ndarray]:
    """:func:`group_csr` with the estimator's own message for a group of one (reference :210)."""
    try:
        return group_csr(index)
    except AssertionError:
        raise AssertionError(f"{what} needs rollout.n > 1.") from None


@torch.no_grad()
def compute_rloo_outcome_advantage(
    token_level_rewards: torch.Tensor, response_mask: torch.Tensor, index: Sequence
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Leave-one-out baseline per uid group. Reference: verl/trainer/core_algos.py:179-214.

    ``a_i = s_i - (sum of the group's other scores) / (n - 1)``, broadcast over the response mask; the same tensor is
    returned as advantages and returns; a group of one sequence is an AssertionError.
    """
    dev, rewards, mask, code = _seq_inputs(token_level_rewards, response_mask)
    bsz, t_len = rewards.shape
    if len(index) != bsz:
        raise ValueError(f"index has {len(index)} entries for a batch of {bsz}")
    order, offsets = group_csr_any(index, "RLOO")
    lib = _lib.load()
    order_d = torch.from_numpy(order).to(dev, non_blocking=True)
    offsets_d = torch.from_numpy(offsets).to(dev, non_blocking=True)
    adv = torch.empty(bsz, t_len, dtype=torch.float32, device=dev)
    seq = torch.empty(2 * max(bsz, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            lib.grpo_rloo_advantage(rewards.data_ptr(), mask.data_ptr(), code, order_d.data_ptr(), offsets_d.data_ptr(),
                                    bsz, t_len, offsets.size - 1, adv.data_ptr(), seq.data_ptr(), _lib.stream_ptr(dev)),
            "grpo_rloo_advantage",
        )
    return adv, adv


@torch.no_grad()
def compute_remax_outcome_advantage(
    token_level_rewards: torch.Tensor, reward_baselines: torch.Tensor, response_mask: torch.Tensor
) -> Tuple[torch.Tensor, torch.Tensor]:
    """ReMax: sequence score minus the greedy-rollout baseline ``(bs,)``. Reference: core_algos.py:248-273."""
    dev, rewards, mask, code = _seq_inputs(token_level_rewards, response_mask)
    require_cuda(reward_baselines)
    bsz, t_len = rewards.shape
    base = f32c(reward_baselines).view(-1)
    if base.numel() != bsz:
        raise ValueError(f"reward_baselines has {base.numel()} entries for a batch of {bsz}")
    lib = _lib.load()
    adv = torch.empty(bsz, t_len, dtype=torch.float32, device=dev)
    seq = torch.empty(2 * max(bsz, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            lib.grpo_remax_advantage(rewards.data_ptr(), base.data_ptr(), mask.data_ptr(), code, bsz, t_len,
                                     adv.data_ptr(), seq.data_ptr(), _lib.stream_ptr(dev)),
            "grpo_remax_advantage",
        )
    return adv, adv


@torch.no_grad()
def compute_reinforce_plus_plus_outcome_advantage(
    token_level_rewards: torch.Tensor, response_mask: torch.Tensor, gamma: float
) -> Tuple[torch.Tensor, torch.Tensor]:
    """REINFORCE++: discounted return-to-go, reset after EOS, whitened over the mask. Reference: core_algos.py:217-245.

    Returns ``(advantages, returns)``.
    """
    dev, rewards, mask, code = _seq_inputs(token_level_rewards, response_mask)
    bsz, t_len = rewards.shape
    lib = _lib.load()
    adv = torch.empty(bsz, t_len, dtype=torch.float32, device=dev)
    ret = torch.empty(bsz, t_len, dtype=torch.float32, device=dev)
    acc = torch.empty(4, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            lib.grpo_reinforce_pp_advantage(rewards.data_ptr(), mask.data_ptr(), code, bsz, t_len, float(gamma),
                                            adv.data_ptr(), ret.data_ptr(), acc.data_ptr(), _lib.stream_ptr(dev)),
            "grpo_reinforce_pp_advantage",
        )
    return adv, ret


def compute_gae_advantage_return(
    token_level_rewards: torch.Tensor, values: torch.Tensor, response_mask: torch.Tensor, gamma: float, lam: float
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Generalised advantage estimation + masked whitening. Reference: core_algos.py:93-133.

    Returns ``(advantages, returns)`` with ``returns = unwhitened advantages + values``.
    """
    dev, rewards, mask, code = _seq_inputs(token_level_rewards, response_mask)
    require_cuda(values)
    if values.shape != rewards.shape:
        raise ValueError("values must have the shape of token_level_rewards")
    bsz, t_len = rewards.shape
    lib = _lib.load()
    vals = f32c(values.detach())
    adv = torch.empty(bsz, t_len, dtype=torch.float32, device=dev)
    ret = torch.empty(bsz, t_len, dtype=torch.float32, device=dev)
    acc = torch.empty(4, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            # `gamma * lam * lastgaelam` (reference :127) multiplies the two Python floats first
            lib.grpo_gae_advantage(rewards.data_ptr(), vals.data_ptr(), mask.data_ptr(), code, bsz, t_len, float(gamma),
                                   float(gamma) * float(lam), adv.data_ptr(), ret.data_ptr(), acc.data_ptr(),
                                   _lib.stream_ptr(dev)),
            "grpo_gae_advantage",
        )
    return adv, ret


def compute_rewards(
    token_level_scores: torch.Tensor, log_probs: torch.Tensor, ref_log_probs: torch.Tensor, kl_ratio: float
) -> torch.Tensor:
    """``token_level_scores - (log_probs - ref_log_probs) * kl_ratio``. Reference: core_algos.py:276-283."""
    dev = require_cuda(token_level_scores, log_probs, ref_log_probs)
    scores, lp, ref = f32c(token_level_scores), f32c(log_probs), f32c(ref_log_probs)
    if lp.shape != scores.shape or ref.shape != scores.shape:
        raise ValueError("token_level_scores, log_probs and ref_log_probs must have the same shape")
    lib = _lib.load()
    out = torch.empty_like(scores)
    cur = torch.empty(1, dtype=torch.float32, device=dev)
    acc = torch.empty(4, dtype=torch.float64, device=dev)
    rows = scores.shape[0] if scores.dim() > 1 else 1
    cols = scores.numel() // max(rows, 1)
    with torch.cuda.device(dev):
        _lib.check(
            lib.grpo_kl_penalty_rewards(scores.data_ptr(), lp.data_ptr(), ref.data_ptr(), None, _lib.MASK_NONE, rows,
                                        cols, _lib.KL_MODES["kl"], float(kl_ratio), out.data_ptr(), cur.data_ptr(),
                                        acc.data_ptr(), _lib.stream_ptr(dev)),
            "grpo_kl_penalty_rewards",
        )
    return out


# ----------------------------------------------------------------------------------------------------------------
# value loss
# ----------------------------------------------------------------------------------------------------------------
class _ValueLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vpreds, returns, values, action_mask, cliprange_value):
        dev = require_cuda(vpreds, returns, values, action_mask)
        lib = _lib.load()
        if returns.shape != vpreds.shape or values.shape != vpreds.shape or action_mask.shape != vpreds.shape:
            raise ValueError("vpreds, returns, values and action_mask must have the same shape")
        vp, ret, old = f32c(vpreds), f32c(returns), f32c(values)
        mask, code = mask_arg(action_mask)
        n = vp.numel()
        dvp = torch.empty(n, dtype=torch.float32, device=dev) if vpreds.requires_grad else None
        out = torch.empty(2, dtype=torch.float32, device=dev)
        acc = torch.empty(4, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(
                lib.grpo_value_loss_fwd_bwd(vp.data_ptr(), ret.data_ptr(), old.data_ptr(), mask.data_ptr(), code, n,
                                            float(cliprange_value), _lib.ptr(dvp), out.data_ptr(), acc.data_ptr(),
                                            _lib.stream_ptr(dev)),
                "grpo_value_loss_fwd_bwd",
            )
        ctx.save_for_backward(dvp)
        ctx.shape, ctx.dtype = vpreds.shape, vpreds.dtype
        vf_loss, vf_clipfrac = out[0], out[1]
        ctx.mark_non_differentiable(vf_clipfrac)
        return vf_loss, vf_clipfrac

    @staticmethod
    def backward(ctx, g_loss, g_clip):
        (dvp,) = ctx.saved_tensors
        grad = None if dvp is None or g_loss is None else (g_loss * dvp.view(ctx.shape)).to(ctx.dtype)
        return grad, None, None, None, None


def compute_value_loss(
    vpreds: torch.Tensor, returns: torch.Tensor, values: torch.Tensor, action_mask: torch.Tensor, cliprange_value: float
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Clipped value loss. Reference: verl/trainer/core_algos.py:356-391.

    Returns 0-d tensors ``(vf_loss, vf_clipfrac)``; ``vf_loss`` is differentiable with respect to ``vpreds``.
    """
    return _ValueLoss.apply(vpreds, returns, values, action_mask, cliprange_value)
