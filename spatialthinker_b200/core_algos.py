"""Drop-in for the GRPO part of ``verl/trainer/core_algos.py``.

``compute_grpo_outcome_advantage`` (reference :137-175), ``compute_policy_loss`` (:291-353) and ``compute_kl`` (:394-436,
also exported as ``kl_penalty``, the upstream-veRL name) keep the reference's signatures, return conventions and
error behaviour; the arithmetic runs in ``csrc/advantage_kernels.cuh`` / ``csrc/loss_kernels.cuh`` through the C ABI.
The other estimators of the reference file (GAE, RLOO, REINFORCE++, ReMax, value loss, KL controllers) are outside
the GRPO path and are not provided.
"""
from __future__ import annotations

from typing import Sequence, Tuple

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
) -> Tuple[torch.Tensor, torch.Tensor]:
    """The same advantages when the batch is sharded by sequence over data-parallel ranks.

    ``token_level_rewards`` / ``response_mask`` are this rank's rows ``[row_begin, row_begin + bs_local)`` of the global
    batch (equal row counts per rank, rank order = row order); ``index_all`` is the uid of EVERY sequence. Scores are
    all-gathered (``bs_all`` floats over NCCL), the group statistics run redundantly on every rank.
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
    scores_all = all_gather_rows(scores, group)
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
        acc = torch.empty(8, dtype=torch.float64, device=dev)
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
