"""The actor's head-side micro-batch loop: drop-in for the hot part of ``verl/workers/actor/dp_actor.py``.

Reference lines mirrored: ``compute_log_prob`` :170-210 (forward-only, micro-batches of
``micro_batch_size_per_device_for_experience``), ``update_policy`` :212-292 (mini-batches of
``global_batch_size_per_device``, GA = global // micro, per-micro-batch loss / GA, metric keys and
``append_to_dict`` list semantics :274-290) and ``_optimizer_step`` :155-167 (clip, skip on non-finite norm).

What differs by design: everything between "final hidden states" and "gradients of hidden states and lm_head weight"
is one fused CUDA pipeline (:mod:`spatialthinker_b200.fused`), the logits tensor never exists, and the 5-6 ``.item()``
host syncs per micro-batch (:274-286) become one device->host copy per ``update_policy`` call.

The transformer body is outside this path. The actor is therefore handed either pre-computed final
hidden states (batch key ``hidden_states``, ``[bs, response_length, H]`` - row ``t`` predicts ``responses[:, t]``,
i.e. the ``[-T-1:-1]`` slice of dp_actor.py:139) or a ``hidden_fn(micro_batch_dict) -> hidden`` callable that runs the
body with autograd enabled; in the second case ``dHidden`` is pushed into that graph with ``hidden.backward(...)``.
"""
from __future__ import annotations

import os
from collections import defaultdict
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional

import torch
import torch.distributed as dist

from . import _lib
from .fused import DeferredDW, compact_index, fused_lm_head_log_probs, gather_rows, grpo_micro_batch_step, scatter_rows
from .sharding import allreduce_mean_

__all__ = ["ActorConfig", "DataParallelPPOActor", "append_to_dict"]


@dataclass
class ActorConfig:
    """The fields of the reference's ``ActorConfig`` (verl/workers/actor/config.py:69-91) that the head path reads."""

    global_batch_size: int = 256
    micro_batch_size_per_device_for_update: int = 4
    micro_batch_size_per_device_for_experience: int = 16
    max_grad_norm: float = 1.0
    clip_ratio_low: float = 0.2
    clip_ratio_high: float = 0.3
    clip_ratio_dual: float = 3.0
    ppo_epochs: int = 1
    # "auto keys" filled by the trainer config's post_init in the reference (verl/trainer/config.py:99-105)
    global_batch_size_per_device: int = -1
    disable_kl: bool = False
    use_kl_loss: bool = False
    kl_penalty: str = "kl"
    kl_coef: float = 0.0
    # extension: 0 in the reference, where the entropy is only logged (dp_actor.py:253)
    entropy_coeff: float = 0.0


def append_to_dict(data: Dict[str, List[Any]], new_data: Dict[str, Any]) -> None:
    """verl/utils/py_functional.py:65."""
    for key, val in new_data.items():
        data.setdefault(key, []).append(val)


def _get(data, name):
    return getattr(data, name) if hasattr(data, name) else data[name]


class DataParallelPPOActor:
    """Head-side PPO/GRPO actor. ``lm_head_weight`` is the bf16 ``[V, H]`` parameter (replicated per rank)."""

    def __init__(
        self,
        config: ActorConfig,
        lm_head_weight: torch.Tensor,
        actor_optimizer: Optional[torch.optim.Optimizer] = None,
        hidden_fn: Optional[Callable[[Dict[str, Any]], torch.Tensor]] = None,
        process_group: Optional["dist.ProcessGroup"] = None,
        compact_padding: bool = True,
        defer_dw: bool = True,
    ):
        self.config = config
        self.rank = int(os.getenv("RANK", "0"))
        self.weight = lm_head_weight
        self.actor_optimizer = actor_optimizer
        self.hidden_fn = hidden_fn
        self.process_group = process_group
        # drop padded token rows before the GEMMs (costs ONE device->host read of the per-micro-batch token counts per
        # update_policy call; the reference computes log-probs for padding and multiplies them by 0)
        self.compact_padding = compact_padding
        # one dW GEMM per group of small micro-batches instead of one per micro-batch (fused.DeferredDW): pays when
        # micro-batches are well below one 18944-row chunk, as with the reference's micro_batch_size_per_device_for_update = 4
        # (+2.6 % tokens/s there, profiles/r1_ab_defer_dw.log); larger micro-batches take the ordinary path unchanged.
        # Costs one more chunk workspace (6 GB at the 7B head).
        self.defer_dw = defer_dw
        self._deferred: Optional[DeferredDW] = None
        self.dweight: Optional[torch.Tensor] = None  # fp32 [V, H] accumulator ("main grad") across micro-batches
        self.last_dhidden: List[torch.Tensor] = []   # per micro-batch dHidden of the last update (when no hidden_fn)

    # ------------------------------------------------------------------------------------------------------------
    def _hidden(self, micro: Dict[str, Any], train: bool) -> torch.Tensor:
        if self.hidden_fn is None:  # an actor built with hidden_fn always runs it (e.g. a reference policy whose
            if "hidden_states" in micro:  # hidden states travel under another key of the same batch)
                return micro["hidden_states"]
            raise KeyError("batch has no 'hidden_states' and the actor was built without hidden_fn")
        if train:
            return self.hidden_fn(micro)
        with torch.no_grad():
            return self.hidden_fn(micro)

    @staticmethod
    def _response_mask(micro: Dict[str, Any]) -> torch.Tensor:
        if "response_mask" in micro:
            return micro["response_mask"]
        response_length = micro["responses"].size(1)
        return micro["attention_mask"][:, -response_length:]  # dp_actor.py:247

    def _optimizer_step(self) -> torch.Tensor:
        """Average dW over ranks, clip by global norm, skip a non-finite step - dp_actor.py:155-167 for the head."""
        assert self.dweight is not None
        allreduce_mean_(self.dweight, self.process_group)
        grad_norm = torch.linalg.vector_norm(self.dweight)
        if self.actor_optimizer is not None:
            clip = torch.clamp(self.config.max_grad_norm / (grad_norm + 1e-6), max=1.0)
            finite = torch.isfinite(grad_norm)
            # a non-finite norm zeroes the step instead of branching on the host (the reference prints and skips)
            scale = torch.where(finite, clip, torch.zeros_like(clip))
            grad = (self.dweight * scale).to(self.weight.dtype)
            grad = torch.where(finite, grad, torch.zeros_like(grad))
            self.weight.grad = grad
            self.actor_optimizer.step()
            self.actor_optimizer.zero_grad()
        self.dweight.zero_()
        return grad_norm

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def compute_log_prob(self, data) -> torch.Tensor:
        """Log-probs of the responses, ``[bs, response_length]`` fp32 - dp_actor.py:170-210 (old / ref log-probs)."""
        temperature = _get(data, "meta_info")["temperature"]
        keys = ["responses"] + [k for k in ("hidden_states", "input_ids", "attention_mask", "position_ids")
                                if k in _get(data, "batch")]
        if self.hidden_fn is not None:  # whatever the body needs travels with the micro-batch
            keys = list(_get(data, "batch").keys())
        if "response_mask" in _get(data, "batch") and "response_mask" not in keys:
            keys.append("response_mask")
        micro_batches = data.select(keys).split(self.config.micro_batch_size_per_device_for_experience)
        # padded response slots are dropped before the GEMM when a mask travels with the batch (ONE device->host read of
        # the per-micro-batch token counts; the reference computes log-probs for padding and masks them later). Padded
        # slots then read 0 instead of the reference's don't-care values.
        counts = None
        if self.compact_padding and micro_batches and (
                "response_mask" in micro_batches[0].batch or "attention_mask" in micro_batches[0].batch):
            sums = [(self._response_mask({**mb.batch}) != 0).sum() for mb in micro_batches]
            counts = torch.stack(sums).cpu().tolist()
        outs = []
        for i, mb in enumerate(micro_batches):
            micro = {**mb.batch, **mb.non_tensor_batch}
            hidden = self._hidden(micro, train=False)
            labels = micro["responses"]
            valid = counts[i] if counts is not None else labels.numel()
            if 0 < valid < labels.numel():
                mask = self._response_mask(micro)
                gather_idx, inverse, _ = compact_index(mask)
                logp_valid, _ = fused_lm_head_log_probs(
                    gather_rows(hidden.reshape(-1, hidden.shape[-1]), gather_idx, valid), self.weight.detach(),
                    gather_rows(labels.reshape(-1), gather_idx, valid), temperature)
                outs.append(scatter_rows(logp_valid, inverse).view(labels.shape))
            else:
                logp, _ = fused_lm_head_log_probs(hidden, self.weight.detach(), labels, temperature)
                outs.append(logp)
        return torch.concat(outs, dim=0)

    def update_policy(self, data) -> Dict[str, Any]:
        """dp_actor.py:212-292: returns the reference's metrics dict (lists per micro-batch / per optimizer step)."""
        cfg = self.config
        temperature = _get(data, "meta_info")["temperature"]  # must be present, as in the reference (:215)
        use_ref = cfg.use_kl_loss and not cfg.disable_kl
        keys = ["responses", "old_log_probs", "advantages"]
        keys += [k for k in ("hidden_states", "input_ids", "attention_mask", "position_ids", "response_mask")
                 if k in _get(data, "batch")]
        if use_ref:
            keys.append("ref_log_probs")
        if self.hidden_fn is not None:
            keys += [k for k in _get(data, "batch").keys() if k not in keys and k != "ref_log_probs"]
        mini_batches = data.select(keys).split(cfg.global_batch_size_per_device)

        if self.dweight is None:
            self.dweight = torch.zeros(self.weight.shape, dtype=torch.float32, device=self.weight.device)
        defer = None
        if self.defer_dw and cfg.entropy_coeff == 0.0:
            if self._deferred is None or self._deferred.dweight is not self.dweight:
                self._deferred = DeferredDW(self.weight.detach(), self.dweight)
            defer = self._deferred
        pending: List[torch.Tensor] = []  # device metric vectors, one per micro-batch
        norms: List[torch.Tensor] = []
        self.last_dhidden = []
        micro_lists = [mini_batch.split(cfg.micro_batch_size_per_device_for_update) for mini_batch in mini_batches]
        counts = None
        if self.compact_padding:
            sums = [(self._response_mask({**mb.batch}) != 0).sum() for mbs in micro_lists for mb in mbs]
            counts = torch.stack(sums).cpu().tolist() if sums else []
        for _ in range(cfg.ppo_epochs):
            flat_i = 0
            for micro_batches in micro_lists:
                grad_accum = cfg.global_batch_size_per_device // cfg.micro_batch_size_per_device_for_update
                for mb in micro_batches:
                    micro = {**mb.batch, **mb.non_tensor_batch}
                    valid_rows = counts[flat_i] if counts is not None else None
                    flat_i += 1
                    hidden = self._hidden(micro, train=True)
                    step = grpo_micro_batch_step(
                        hidden.detach(), self.weight.detach(), micro["responses"], micro["old_log_probs"],
                        micro["advantages"], micro["ref_log_probs"] if use_ref else None, self._response_mask(micro),
                        temperature=temperature, clip_ratio_low=cfg.clip_ratio_low, clip_ratio_high=cfg.clip_ratio_high,
                        clip_ratio_dual=cfg.clip_ratio_dual, kl_penalty=cfg.kl_penalty if use_ref else None,
                        kl_coef=cfg.kl_coef, grad_accum=float(grad_accum), entropy_coeff=cfg.entropy_coeff,
                        dweight_accum=self.dweight, valid_rows=valid_rows, defer=defer,
                    )
                    if hidden.requires_grad:
                        hidden.backward(step["dhidden"])  # continue into the transformer body
                    else:
                        self.last_dhidden.append(step["dhidden"])
                    pending.append(step["metrics"])
                if defer is not None:
                    defer.flush()  # dW complete before it is all-reduced and applied
                norms.append(self._optimizer_step())

        # one device->host transfer for every scalar of this call
        host = torch.stack(pending).float().cpu() if pending else torch.zeros(0, _lib.NUM_METRICS)
        host_norms = torch.stack(norms).float().cpu().tolist() if norms else []
        metrics: Dict[str, Any] = defaultdict(list)
        per_step = len(pending) // max(len(norms), 1)
        for i, row in enumerate(host.tolist()):
            if use_ref:  # the reference ASSIGNS these two (dp_actor.py:274-275): last micro-batch wins
                metrics["actor/kl_loss"] = row[_lib.MET_KL_LOSS]
                metrics["actor/kl_coef"] = cfg.kl_coef
            append_to_dict(metrics, {
                "actor/pg_loss": row[_lib.MET_TOTAL],
                "actor/pg_clipfrac_higher": row[_lib.MET_CLIPFRAC_HI],
                "actor/pg_clipfrac_lower": row[_lib.MET_CLIPFRAC_LO],
                "actor/entropy_loss": row[_lib.MET_ENTROPY],
                "actor/ppo_kl": row[_lib.MET_PPO_KL],
            })
            if per_step and (i + 1) % per_step == 0:
                append_to_dict(metrics, {"actor/grad_norm": host_norms[(i + 1) // per_step - 1]})
        return metrics
