"""The two driver-side data steps of ``verl/trainer/ray_trainer.py`` that feed the actor update: KL reward shaping
(``apply_kl_penalty`` :125-145) and advantage estimation (``compute_advantage`` :148-175). SURVEY.md §8 rows a11, f-3.

Nothing of Ray lives here - the module keeps the reference's name so that ``patch.py`` can swap the two functions into
a live checkout. ``data`` is any object with ``.batch`` (name -> tensor) and ``.non_tensor_batch`` mappings: the real
``DataProto`` or :class:`spatialthinker_b200.protocol.TensorBatch`. Tensors must be on the GPU; in the reference these
steps run on the driver's CPU with Python loops over the batch (0.21 s at 4096 sequences) and a
``.item()`` per call - here each is a handful of kernel launches on the stream that already holds the log-probs, and
the one scalar the KL controller needs is the only device->host read.
"""
from __future__ import annotations

from enum import Enum
from typing import Any, Dict, Tuple

import torch

from . import _lib, core_algos
from ._util import f32c, mask_arg, require_cuda

__all__ = ["AdvantageEstimator", "apply_kl_penalty", "compute_advantage", "experience_pass", "kl_penalty_rewards"]


class AdvantageEstimator(str, Enum):
    """ray_trainer.py:66-75."""

    GAE = "gae"
    GRPO = "grpo"
    REINFORCE_PLUS_PLUS = "reinforce_plus_plus"
    REMAX = "remax"
    RLOO = "rloo"


def _batch_size(data) -> int:
    size = getattr(data.batch, "batch_size", None)  # TensorDict
    if size is not None:
        return int(size[0])
    return len(data)


def kl_penalty_rewards(
    token_level_scores: torch.Tensor, old_log_probs: torch.Tensor, ref_log_probs, response_mask: torch.Tensor,
    kl_coef: float, kl_penalty: str = "kl",
) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(token_level_rewards, current_kl)`` of ray_trainer.py:131-142 in one kernel: ``scores - kl_coef * kld`` with
    ``kld = compute_kl(old, ref) * mask`` (zeros without a reference policy), and the batch mean of the per-sequence
    masked means of ``kld`` as a 1-element device tensor."""
    dev = require_cuda(token_level_scores, response_mask)
    scores = f32c(token_level_scores)
    if scores.dim() != 2 or response_mask.shape != scores.shape:
        raise ValueError("token_level_scores and response_mask must both be (bs, response_length)")
    mask, code = mask_arg(response_mask)
    mode = _lib.KL_MODES.get(kl_penalty) if isinstance(kl_penalty, str) else None
    lp = ref = None
    if ref_log_probs is not None:
        if mode is None or mode < 0:
            raise NotImplementedError(f"Unknown KL penalty: {kl_penalty}.")
        require_cuda(old_log_probs, ref_log_probs)
        lp, ref = f32c(old_log_probs), f32c(ref_log_probs)
        if lp.shape != scores.shape or ref.shape != scores.shape:
            raise ValueError("log-probs must have the shape of token_level_scores")
    bsz, t_len = scores.shape
    rewards = torch.empty_like(scores)
    current = torch.zeros(1, dtype=torch.float32, device=dev)
    acc = torch.empty(4, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            _lib.load().grpo_kl_penalty_rewards(scores.data_ptr(), _lib.ptr(lp), _lib.ptr(ref), mask.data_ptr(), code,
                                                bsz, t_len, mode if ref is not None else -1, float(kl_coef),
                                                rewards.data_ptr(), current.data_ptr(), acc.data_ptr(),
                                                _lib.stream_ptr(dev)),
            "grpo_kl_penalty_rewards",
        )
    return rewards, current


def apply_kl_penalty(data, kl_ctrl: "core_algos.KLController", kl_penalty: str = "kl"):
    """ray_trainer.py:125-145: writes ``token_level_rewards`` into the batch, updates the controller with the measured
    KL and the batch size, returns ``(data, {"critic/kl", "critic/kl_coef"})``."""
    batch = data.batch
    has_ref = "ref_log_probs" in batch.keys()
    rewards, current = kl_penalty_rewards(
        batch["token_level_scores"], batch["old_log_probs"] if has_ref else None,
        batch["ref_log_probs"] if has_ref else None, batch["response_mask"], kl_ctrl.kl_coef, kl_penalty)
    batch["token_level_rewards"] = rewards
    current_kl = float(current.item())
    metrics = {"critic/kl": current_kl, "critic/kl_coef": kl_ctrl.kl_coef}
    kl_ctrl.update(current_kl=current_kl, n_steps=_batch_size(data))
    return data, metrics


def compute_advantage(data, adv_estimator, gamma: float = 1.0, lam: float = 1.0):
    """ray_trainer.py:148-175: dispatch on the estimator, store ``advantages`` and ``returns`` in the batch."""
    batch: Dict[str, Any] = data.batch
    rewards, mask = batch["token_level_rewards"], batch["response_mask"]
    if adv_estimator == AdvantageEstimator.GAE:
        adv, ret = core_algos.compute_gae_advantage_return(rewards, batch["values"], mask, gamma, lam)
    elif adv_estimator == AdvantageEstimator.GRPO:
        adv, ret = core_algos.compute_grpo_outcome_advantage(rewards, mask, data.non_tensor_batch["uid"])
    elif adv_estimator == AdvantageEstimator.REINFORCE_PLUS_PLUS:
        adv, ret = core_algos.compute_reinforce_plus_plus_outcome_advantage(rewards, mask, gamma)
    elif adv_estimator == AdvantageEstimator.REMAX:
        adv, ret = core_algos.compute_remax_outcome_advantage(rewards, batch["reward_baselines"], mask)
    elif adv_estimator == AdvantageEstimator.RLOO:
        adv, ret = core_algos.compute_rloo_outcome_advantage(rewards, mask, data.non_tensor_batch["uid"])
    else:
        raise NotImplementedError
    batch["advantages"] = adv
    batch["returns"] = ret
    return data


def experience_pass(data, actor, ref_actor=None, *, adv_estimator=AdvantageEstimator.GRPO, use_kl_loss: bool = True,
                    kl_ctrl=None, kl_penalty: str = "kl", gamma: float = 1.0, lam: float = 1.0):
    """The head-side experience pass of ``RayPPOTrainer.fit`` between rollout and ``update_actor``
    (ray_trainer.py:633-663): old log-probs from the actor, reference log-probs from the reference policy, KL reward
    shaping (only when the KL is NOT a loss term, :650-656, else rewards = scores), advantages.

    ``actor`` / ``ref_actor`` are :class:`spatialthinker_b200.dp_actor.DataParallelPPOActor` objects (the reference
    reaches them through Ray worker groups); each contributes one forward-only sweep of the fused head (2*H*V FLOP per
    token, logits never materialised). Everything stays on the device; the only host read is the scalar the KL
    controller consumes. Returns ``(data, metrics)``.
    """
    batch = data.batch
    metrics: Dict[str, Any] = {}
    batch["old_log_probs"] = actor.compute_log_prob(data)                    # :634-636
    if ref_actor is not None:
        batch["ref_log_probs"] = ref_actor.compute_log_prob(data)            # :639-642
    if not use_kl_loss and ref_actor is not None:
        if kl_ctrl is None:
            raise ValueError("KL reward shaping needs a KL controller (core_algos.get_kl_controller)")
        data, kl_metrics = apply_kl_penalty(data, kl_ctrl=kl_ctrl, kl_penalty=kl_penalty)
        metrics.update(kl_metrics)
    else:
        batch["token_level_rewards"] = batch["token_level_scores"]          # :658
    data = compute_advantage(data, adv_estimator=adv_estimator, gamma=gamma, lam=lam)
    return data, metrics
