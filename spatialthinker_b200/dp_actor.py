"""The actor's head-side micro-batch loop: drop-in for the hot part of ``verl/workers/actor/dp_actor.py``.

Reference lines mirrored: ``compute_log_prob`` :170-210 (forward-only, micro-batches of
``micro_batch_size_per_device_for_experience``), ``update_policy`` :212-292 (mini-batches of
``global_batch_size_per_device``, GA = global // micro, per-micro-batch loss / GA, metric keys and
``append_to_dict`` list semantics :274-290) and ``_optimizer_step`` :155-167 (global-norm clip over every parameter the
optimizer holds, skip on a non-finite norm).

What differs by design: everything between "final hidden states" and "gradients of hidden states and lm_head weight"
is one fused CUDA pipeline (:mod:`spatialthinker_b200.fused`), the logits tensor never exists, and the 5-6 ``.item()``
host syncs per micro-batch (:274-286) become one device->host copy per ``update_policy`` call (plus the reference's own
``isfinite`` branch per optimizer step when an optimizer is attached).

The transformer body is outside this path. The actor is therefore handed either pre-computed final
hidden states (batch key ``hidden_states``, ``[bs, response_length, H]`` - row ``t`` predicts ``responses[:, t]``,
i.e. the ``[-T-1:-1]`` slice of dp_actor.py:139) or a ``hidden_fn(micro_batch_dict) -> hidden`` callable that runs the
body with autograd enabled; in the second case ``dHidden`` is pushed into that graph with ``hidden.backward(...)``.

Token-balanced micro-batching (``ActorConfig.use_dynamic_bsz``): the reference carries
``rearrange_micro_batches`` (verl/utils/seqlen_balancing.py:222-255) - split a mini-batch into
``ceil(tokens / max_token_len)`` micro-batches with Karmarkar-Karp-balanced token sums - without wiring it into its
actor; upstream veRL does (``use_dynamic_bsz``), weighting each micro-batch's loss by its share of the mini-batch's
sequences. The same here: with ragged responses every micro-batch then fills whole 18 944-row chunks of the GEMM
pipeline after padding has been compacted away.
"""
from __future__ import annotations

import os
import warnings

import numpy as np
from collections import defaultdict
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib
from .fused import DeferredDW, compact_index, fused_lm_head_log_probs, gather_rows, grpo_micro_batch_step, scatter_rows
from .protocol import TensorBatch
from .sharding import allreduce_mean_, micro_batch_counts, rearrange_micro_batches

__all__ = ["ActorConfig", "DataParallelPPOActor", "append_to_dict", "grad_sumsq", "grad_scale_cast"]


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
    # extension (upstream veRL's use_dynamic_bsz / ppo_max_token_len_per_gpu; the reference only ships the helper,
    # seqlen_balancing.py:222-255): micro-batches are formed by token count instead of sequence count. Each micro-batch's
    # loss is weighted by len(micro) / len(mini) - upstream's convention - instead of 1 / GA.
    use_dynamic_bsz: bool = False
    max_token_len_per_micro_batch: int = 37888  # two 18 944-row chunks of the GEMM pipeline
    # extension (speed-aware shards, sharding.speed_weighted_counts): when the ranks hold DIFFERENT numbers of sequences per
    # mini-batch, every rank scales its micro-batch losses for the same nominal mini-batch size (global batch / world size)
    # so that the mean all-reduce over ranks still weights every sequence alike. 0: global_batch_size_per_device.
    loss_scale_batch_size: int = 0
    # extension (config C4 of BASELINE.json): also compute the true per-token entropy lse - sum p z in the same pass and
    # log its masked mean as actor/entropy (the reference only logs the estimator -masked_mean(log p), dp_actor.py:253)
    log_true_entropy: bool = False
    # extension: run-to-run bit-reproducible gradients (library option "deterministic": no split-K, ordered one-hot rows).
    # The option is process-wide; it is applied when the actor is built.
    deterministic: bool = False


def append_to_dict(data: Dict[str, List[Any]], new_data: Dict[str, Any]) -> None:
    """verl/utils/py_functional.py:65."""
    for key, val in new_data.items():
        data.setdefault(key, []).append(val)


def _get(data, name):
    return getattr(data, name) if hasattr(data, name) else data[name]


# ----------------------------------------------------------------------------------------------------------------
# passes over the fp32 gradient accumulator (csrc/grad_kernels.cuh)
# ----------------------------------------------------------------------------------------------------------------
def grad_sumsq(grad: torch.Tensor, zero_after: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``sum(grad ** 2)`` of a contiguous fp32 CUDA tensor as a 1-element float64 device tensor (``out`` given: added to
    it). ``zero_after`` zeroes ``grad`` in the same pass. Partial sums are reduced in a fixed order."""
    if not grad.is_cuda or grad.dtype != torch.float32 or not grad.is_contiguous():
        raise ValueError("grad must be a contiguous float32 CUDA tensor")
    dev = grad.device
    acc = out if out is not None else torch.empty(1, dtype=torch.float64, device=dev)
    tmp = torch.empty(_lib.GRAD_SCRATCH_DOUBLES, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().grpo_grad_sumsq(grad.data_ptr(), grad.numel(), int(zero_after), int(out is not None),
                                               acc.data_ptr(), tmp.data_ptr(), _lib.stream_ptr(dev)), "grpo_grad_sumsq")
    return acc


def grad_scale_cast(grad: torch.Tensor, scale, out: torch.Tensor, zero_after: bool = False) -> torch.Tensor:
    """``out = bf16(grad * scale)`` (``scale``: python float or 1-element fp32 device tensor); ``zero_after`` zeroes
    ``grad`` in the same pass."""
    if not grad.is_cuda or grad.dtype != torch.float32 or not grad.is_contiguous():
        raise ValueError("grad must be a contiguous float32 CUDA tensor")
    if out.dtype != torch.bfloat16 or out.shape != grad.shape or not out.is_contiguous() or out.device != grad.device:
        raise ValueError("out must be a contiguous bfloat16 tensor shaped like grad on the same device")
    dev = grad.device
    sdev = scale if isinstance(scale, torch.Tensor) else None
    if sdev is not None and (sdev.dtype != torch.float32 or sdev.numel() != 1 or sdev.device != dev):
        raise ValueError("a device scale must be one float32 element on grad's device")
    with torch.cuda.device(dev):
        _lib.check(_lib.load().grpo_grad_scale_cast(grad.data_ptr(), grad.numel(), _lib.ptr(sdev),
                                                    1.0 if sdev is not None else float(scale), out.data_ptr(),
                                                    int(zero_after), _lib.stream_ptr(dev)), "grpo_grad_scale_cast")
    return out


def _select_rows(mb, idx: Sequence[int]):
    """Rows ``idx`` of a batch object as a :class:`TensorBatch` (``TensorBatch.take`` / a TensorDict-backed DataProto)."""
    if hasattr(mb, "take"):
        return mb.take(idx)
    index = torch.as_tensor(list(idx), dtype=torch.long)
    batch = {k: v[index.to(v.device)] for k, v in mb.batch.items()}
    non_tensor = {k: v[index.numpy()] for k, v in getattr(mb, "non_tensor_batch", {}).items()}
    return TensorBatch(batch, non_tensor, getattr(mb, "meta_info", {}))


def _rows_of(t: torch.Tensor, rows: Sequence[int]) -> torch.Tensor:
    """Batch rows ``rows`` of a device tensor: a view when they are one consecutive run (the reference's ``split``),
    else a gathered copy (token-balanced micro-batches)."""
    if len(rows) > 0 and rows[-1] - rows[0] + 1 == len(rows) and all(b - a == 1 for a, b in zip(rows, rows[1:])):
        return t[rows[0]:rows[-1] + 1]
    return t.index_select(0, torch.as_tensor(list(rows), dtype=torch.long, device=t.device))


class _HostStager:
    """Pinned-host -> device staging of micro-batches on a copy stream, two slots: the copy of the next micro-batch
    runs while the current one computes. Rows are copied as consecutive runs straight out of the caller's host tensors
    (no host-side gather)."""

    def __init__(self, device: torch.device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.bufs: List[Dict[str, torch.Tensor]] = [{}, {}]
        self.views: List[Dict[str, torch.Tensor]] = [{}, {}]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]
        self.h2d_bytes = 0

    def prefetch(self, slot: int, batch: Dict[str, torch.Tensor], rows: Sequence[int]) -> None:
        runs = []  # [first, last + 1) runs of consecutive rows
        for r in rows:
            if runs and runs[-1][1] == r:
                runs[-1][1] = r + 1
            else:
                runs.append([r, r + 1])
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.free[slot])  # the kernels that read this slot last have finished
            for key, src in batch.items():
                shape = (len(rows),) + tuple(src.shape[1:])
                buf = self.bufs[slot].get(key)
                if buf is None or buf.dtype != src.dtype or buf.shape[1:] != shape[1:] or buf.shape[0] < shape[0]:
                    buf = torch.empty(shape, dtype=src.dtype, device=self.device)
                    self.bufs[slot][key] = buf
                dst = buf[:len(rows)]
                at = 0
                for a, b in runs:
                    dst[at:at + b - a].copy_(src[a:b], non_blocking=True)
                    at += b - a
                self.views[slot][key] = dst
                self.h2d_bytes += dst.numel() * dst.element_size()
            self.ready[slot].record(self.stream)

    def get(self, slot: int) -> Dict[str, torch.Tensor]:
        torch.cuda.current_stream(self.device).wait_event(self.ready[slot])
        return dict(self.views[slot])

    def release(self, slot: int) -> None:
        self.free[slot].record(torch.cuda.current_stream(self.device))


class DataParallelPPOActor:
    """Head-side PPO/GRPO actor. ``lm_head_weight`` is the bf16 ``[V, H]`` parameter (replicated per rank).

    ``actor_optimizer`` (optional) holds the parameters to be stepped: the head weight and, when ``hidden_fn`` runs a
    body with autograd, the body's. ``_optimizer_step`` clips by the GLOBAL norm over all of them like the reference's
    ``clip_grad_norm_(actor_module.parameters())``. Gradients of body parameters are averaged over ranks by whatever wraps
    the body (DDP), or here when ``reduce_body_grads=True`` (a bare replicated body); an FSDP-sharded body keeps its own
    clipping and belongs to the autograd-level integration (level 2), not to this class.
    """

    def __init__(
        self,
        config: ActorConfig,
        lm_head_weight: torch.Tensor,
        actor_optimizer: Optional[torch.optim.Optimizer] = None,
        hidden_fn: Optional[Callable[[Dict[str, Any]], torch.Tensor]] = None,
        process_group: Optional["dist.ProcessGroup"] = None,
        compact_padding: bool = True,
        defer_dw: bool = True,
        reduce_body_grads: Optional[bool] = None,
        peer_exchange: Optional[bool] = None,
    ):
        self.config = config
        self.rank = int(os.getenv("RANK", "0"))
        self.weight = lm_head_weight
        self.actor_optimizer = actor_optimizer
        self.hidden_fn = hidden_fn
        self.process_group = process_group
        # drop padded token rows before the GEMMs (costs ONE device->host read of the per-sequence token counts per
        # update_policy call; the reference computes log-probs for padding and multiplies them by 0)
        self.compact_padding = compact_padding
        # one dW GEMM per group of small micro-batches instead of one per micro-batch (fused.DeferredDW): pays when
        # micro-batches are well below one 18944-row chunk, as with the reference's micro_batch_size_per_device_for_update = 4
        # (+2.6 % tokens/s there, profiles/r1_ab_defer_dw.log); larger micro-batches take the ordinary path unchanged.
        # Costs one more chunk workspace (6 GB at the 7B head), allocated when the first small micro-batch arrives.
        self.defer_dw = defer_dw
        self.reduce_body_grads = reduce_body_grads
        # dW exchange over peer-mapped memory (peer.PeerGroup: reduce-scatter + norm, then clip + bf16 all-gather + zero, this
        # library's kernels over NVLink) instead of NCCL's fp32 all-reduce followed by separate norm / clip / cast passes.
        # None: on when the ranks of the group share one host (GRPO_PEER_EXCHANGE=0 turns it off); if the GPUs cannot map
        # each other's memory every rank falls back to NCCL together, with a warning. True: failure to set it up raises.
        self.peer_exchange = peer_exchange
        self._peer = None          # None: not tried yet, False: unavailable, else (PeerGroup, dW buffer, gradient buffer)
        self._deferred: Optional[DeferredDW] = None
        self.dweight: Optional[torch.Tensor] = None  # fp32 [V, H] accumulator ("main grad") across micro-batches
        self._grad_buf: Optional[torch.Tensor] = None  # bf16 [V, H]: what the optimizer sees as weight.grad
        self._stager: Optional[_HostStager] = None     # staging buffers of host-resident batches (kept across calls)
        self._warned_body = False
        # measurement aid (bench.py): CUDA events around every dW all-reduce; the time the slowest-arriving rank spends in
        # it is the wire time, what the others spend on top of that is waiting for it
        self.time_collectives = False
        self.collective_events: List[tuple] = []
        self.last_dhidden: List[torch.Tensor] = []   # per micro-batch dHidden of the last update (when no hidden_fn)
        if config.deterministic:
            _lib.check(_lib.load().grpo_set_option(b"deterministic", 1), "grpo_set_option")

    def release_workspaces(self) -> None:
        """Give back the deferred-dW chunk workspace, the fp32 accumulator and the gradient buffer (e.g. before the
        rollout engine needs the memory); they are rebuilt on the next ``update_policy``. With the peer exchange on this
        is COLLECTIVE over the group: the other ranks have the accumulator and the gradient buffer mapped and must unmap
        them before the memory may be freed."""
        if self._deferred is not None:
            self._deferred.release()
        self._deferred = None
        if self._peer:  # collective: the other ranks have these buffers mapped
            group, dw_buf, grad_buf = self._peer
            group.release(dw_buf)
            group.release(grad_buf)
        self._peer = None
        self.dweight = None
        self._grad_buf = None
        self._stager = None

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

    def _other_params(self) -> List[torch.nn.Parameter]:
        """Optimizer parameters other than the head weight that carry a gradient (the body behind ``hidden_fn``)."""
        if self.actor_optimizer is None:
            return []
        out = []
        for group in self.actor_optimizer.param_groups:
            for p in group["params"]:
                if p is not self.weight and p.grad is not None:
                    out.append(p)
        return out

    def _setup_peer(self) -> None:
        """Map the fp32 accumulator and the bf16 gradient buffer into every rank of the group (collective, once)."""
        if self._peer is not None:
            return
        self._peer = False
        world = dist.get_world_size(self.process_group) if dist.is_available() and dist.is_initialized() else 1
        want = self.peer_exchange
        if want is None:
            want = os.environ.get("GRPO_PEER_EXCHANGE", "1") != "0"
        if world == 1 or not want or not self.weight.is_cuda or self.dweight.numel() % 8 != 0:
            return
        from . import peer

        try:
            group = peer.get_group(self.process_group, self.weight.device)
            if self._grad_buf is None or self._grad_buf.shape != self.weight.shape:
                self._grad_buf = torch.empty(self.weight.shape, dtype=torch.bfloat16, device=self.weight.device)
            self._peer = (group, group.register(self.dweight), group.register(self._grad_buf))
        except peer.PeerUnavailable as exc:
            if self.peer_exchange:
                raise
            if self.rank == 0:
                warnings.warn(f"DataParallelPPOActor: peer-mapped dW exchange unavailable ({exc}); using NCCL's all-reduce.")

    def _timed(self, fn):
        """Run ``fn`` between CUDA events when the bench asks for per-rank collective times."""
        if not self.time_collectives:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.collective_events.append((e0, e1))
        return out

    def _optimizer_step_peer(self) -> torch.Tensor:
        """``_optimizer_step`` over peer-mapped memory: the same arithmetic - fp32 mean over ranks, global norm, clip,
        bf16 gradient for the optimizer, non-finite skip - with the exchange and the passes over ``dW`` fused
        (csrc/peer_kernels.cuh). Every rank issues the same kernels whatever the norm turns out to be."""
        group, dw_buf, grad_buf = self._peer
        cfg = self.config
        opt = self.actor_optimizer
        others = self._other_params()
        reduce_body = bool(self.reduce_body_grads)
        wgrad = self.weight.grad if (opt is not None and self.weight.is_leaf) else None
        if wgrad is not None and wgrad is not self._grad_buf:
            # a tied embedding's share of this parameter's gradient joins the accumulator BEFORE the exchange: it is
            # averaged with it (already-averaged values are a fixed point of the mean)
            self.dweight.add_(wgrad)
        if reduce_body:
            for p in others:
                allreduce_mean_(p.grad, self.process_group)
        total = self._timed(lambda: group.reduce_scatter_sumsq(dw_buf))
        if others:
            norms = torch._foreach_norm([p.grad for p in others])
            total = total + torch.stack([n.double() for n in norms]).square().sum()
        grad_norm = total.sqrt().float().squeeze(0)
        clip = torch.clamp(cfg.max_grad_norm / (grad_norm + 1e-6), max=1.0).reshape(1)
        self._timed(lambda: group.scale_cast_allgather(dw_buf, grad_buf, clip, zero_after=True))
        if opt is None:  # head-only accumulation (bench, tests): the clipped bf16 gradient stays in self._grad_buf
            return grad_norm
        if not bool(torch.isfinite(grad_norm)):
            print("Gradient norm is not finite. Skip update.")
        else:
            self.weight.grad = self._grad_buf
            for p in others:
                p.grad.mul_(clip.reshape(()).to(p.grad.dtype))
            opt.step()
        opt.zero_grad()
        if self.weight.is_leaf:
            self.weight.grad = None
        return grad_norm

    def _optimizer_step(self) -> torch.Tensor:
        """dp_actor.py:155-167. The head's ``dW`` is averaged over ranks in fp32 (FSDP's ``mp_reduce_dtype``,
        actor/config.py:58); the clip coefficient comes from the global norm over EVERY parameter the optimizer holds;
        a non-finite norm skips the update (the reference's host branch) and the gradients are dropped."""
        assert self.dweight is not None
        if self._peer:
            return self._optimizer_step_peer()
        cfg = self.config
        if self.time_collectives:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            allreduce_mean_(self.dweight, self.process_group)
            e1.record()
            self.collective_events.append((e0, e1))
        else:
            allreduce_mean_(self.dweight, self.process_group)
        opt = self.actor_optimizer
        if opt is None:  # head-only accumulation (bench, tests): norm and zeroing share one pass over dW
            return grad_sumsq(self.dweight, zero_after=True).sqrt().float().squeeze(0)
        world = dist.get_world_size(self.process_group) if dist.is_available() and dist.is_initialized() else 1
        others = self._other_params()
        reduce_body = bool(self.reduce_body_grads)
        if others and world > 1 and self.reduce_body_grads is None and not self._warned_body:
            warnings.warn("DataParallelPPOActor: the optimizer holds body parameters and reduce_body_grads was not given; "
                          "their gradients are assumed to be averaged by the wrapper that owns the body (DDP).")
            self._warned_body = True
        # whatever the body's backward left on the head weight itself (3B checkpoints tie lm_head to embed_tokens) is
        # part of the same parameter's gradient: fold it into the fp32 accumulator instead of overwriting it
        wgrad = self.weight.grad if self.weight.is_leaf else None
        if wgrad is not None:
            if reduce_body:
                allreduce_mean_(wgrad, self.process_group)
            self.dweight.add_(wgrad)
        if reduce_body:
            for p in others:
                allreduce_mean_(p.grad, self.process_group)
        total = grad_sumsq(self.dweight)
        if others:
            norms = torch._foreach_norm([p.grad for p in others])
            total = total + torch.stack([n.double() for n in norms]).square().sum()
        grad_norm = total.sqrt().float().squeeze(0)
        clip = torch.clamp(cfg.max_grad_norm / (grad_norm + 1e-6), max=1.0).reshape(1)  # clip_grad_norm_'s coefficient
        if not bool(torch.isfinite(grad_norm)):  # one host read per optimizer step, as in the reference (:161)
            print("Gradient norm is not finite. Skip update.")
            self.dweight.zero_()
        else:
            if self._grad_buf is None or self._grad_buf.shape != self.weight.shape:
                self._grad_buf = torch.empty(self.weight.shape, dtype=torch.bfloat16, device=self.weight.device)
            grad_scale_cast(self.dweight, clip, self._grad_buf, zero_after=True)  # clip + cast + zero: one pass
            self.weight.grad = self._grad_buf
            for p in others:
                p.grad.mul_(clip.reshape(()).to(p.grad.dtype))
            opt.step()
        opt.zero_grad()
        if self.weight.is_leaf:
            self.weight.grad = None  # the bf16 buffer is ours; a tied embedding's next backward starts from nothing
        return grad_norm

    # ------------------------------------------------------------------------------------------------------------
    def _valid_lengths(self, data) -> Optional[List[int]]:
        """Per-sequence count of unmasked response slots, read to the host ONCE per call (None: no mask in the batch)."""
        batch = _get(data, "batch")
        if "response_mask" in batch:
            mask = batch["response_mask"]
        elif "attention_mask" in batch and "responses" in batch:
            mask = batch["attention_mask"][:, -batch["responses"].size(1):]
        else:
            return None
        return (mask != 0).sum(dim=1).cpu().tolist()

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
        micro_size = self.config.micro_batch_size_per_device_for_experience
        micro_batches = data.select(keys).split(micro_size)
        # padded response slots are dropped before the GEMM when a mask travels with the batch (ONE device->host read of
        # the per-sequence token counts; the reference computes log-probs for padding and masks them later). Padded
        # slots then read 0 instead of the reference's don't-care values.
        lens = self._valid_lengths(data) if self.compact_padding else None
        outs = []
        row = 0
        for mb in micro_batches:
            micro = {**mb.batch, **mb.non_tensor_batch}
            hidden = self._hidden(micro, train=False)
            labels = micro["responses"]
            valid = sum(lens[row:row + labels.shape[0]]) if lens is not None else labels.numel()
            row += labels.shape[0]
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

    def _micro_plan(self, rows: List[int], lens: Optional[List[int]], t_len: int, num_micro: Optional[int] = None):
        """[(row ids, loss divisor, valid token rows or None)] for the mini-batch made of batch rows ``rows``.

        Fixed size (the reference, dp_actor.py:233-237): consecutive runs of ``micro_batch_size_per_device_for_update``
        rows, divisor GA. ``use_dynamic_bsz``: ``rearrange_micro_batches`` by token count (seqlen_balancing.py:222-255),
        divisor ``len(mini) / len(micro)``."""
        cfg = self.config
        n = len(rows)
        nominal = cfg.loss_scale_batch_size or n
        count = (lambda idx: None) if lens is None else (lambda idx: sum(lens[i] for i in idx))
        if cfg.use_dynamic_bsz:
            tokens = [lens[i] for i in rows] if lens is not None else [t_len] * n
            cap = max(cfg.max_token_len_per_micro_batch, t_len)
            if min(tokens) == max(tokens):
                # equal lengths (dense responses): every split into equal counts is perfectly balanced, so consecutive runs
                # do - no partitioning work on the host and no row gather on the device (views of the batch)
                num = num_micro if num_micro is not None else micro_batch_counts([sum(tokens)], cap, self.process_group,
                                                                                 self.weight.device)[0]
                base, extra = divmod(n, num)
                edges = [0]
                for j in range(num):
                    edges.append(edges[-1] + base + (1 if j < extra else 0))
                parts = [list(range(edges[j], edges[j + 1])) for j in range(num)]
            else:
                parts = rearrange_micro_batches(tokens, cap, self.process_group, device=self.weight.device,
                                                num_micro_batches=num_micro)
            return [([rows[i] for i in p], nominal / len(p), count([rows[i] for i in p])) for p in parts]
        micro = cfg.micro_batch_size_per_device_for_update
        # the reference's split() only knows equal chunks (protocol.py:488-523); a speed-aware shard (loss_scale_batch_size
        # set) may end in a shorter micro-batch, weighted by its own length like every other
        assert n % micro == 0 or cfg.loss_scale_batch_size > 0, (
            f"only support equal chunk. Got size of DataProto {n} and chunk {n // max(micro, 1)}.")
        chunks = [rows[m0:m0 + micro] for m0 in range(0, n, micro)]
        return [(c, nominal / len(c), count(c)) for c in chunks]  # full chunks: nominal / micro = the reference's GA (:233)

    def update_policy(self, data) -> Dict[str, Any]:
        """dp_actor.py:212-292: returns the reference's metrics dict (lists per micro-batch / per optimizer step).

        The batch may live on the device (the reference's workers call ``data.to("cuda")`` first, fsdp_workers.py) or in
        (pinned) host memory: host batches are streamed micro-batch by micro-batch on a copy stream, double-buffered,
        so the copy of micro-batch i + 1 overlaps the GEMMs of micro-batch i."""
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
        selected = data.select(keys)
        batch = {k: selected.batch[k] for k in keys}
        non_tensor = dict(getattr(selected, "non_tensor_batch", {}))
        n_total = batch["responses"].shape[0]
        t_len = int(batch["responses"].shape[1])
        mini = cfg.global_batch_size_per_device
        assert mini > 0 and n_total % mini == 0, (
            f"only support equal chunk. Got size of DataProto {n_total} and chunk {n_total // max(mini, 1)}.")
        lens = self._valid_lengths(selected) if (self.compact_padding or cfg.use_dynamic_bsz) else None
        on_host = not batch["responses"].is_cuda
        if any(v.is_cuda == on_host for v in batch.values()):
            raise ValueError("update_policy: the batch must live entirely on the device or entirely in host memory")
        stager = None
        if on_host:
            if self._stager is None:
                self._stager = _HostStager(self.weight.device)
            stager = self._stager

        if self.dweight is None:
            self.dweight = torch.zeros(self.weight.shape, dtype=torch.float32, device=self.weight.device)
        self._setup_peer()
        defer = None
        if self.defer_dw and cfg.entropy_coeff == 0.0:
            if self._deferred is None or self._deferred.dweight is not self.dweight:
                self._deferred = DeferredDW(self.weight.detach(), self.dweight)
            defer = self._deferred
        pending: List[torch.Tensor] = []  # device metric vectors, one per micro-batch
        norms: List[torch.Tensor] = []
        step_of: List[int] = []           # optimizer step each micro-batch belongs to
        self.last_dhidden = []
        # mini-batches in order; the micro-batch plan of mini-batch i + 1 is worked out on the host right after the first
        # micro-batch of mini-batch i has been enqueued (Karmarkar-Karp over 1024 sequences takes ~30 ms: hidden behind
        # the GPU's queue instead of in front of it). With token-balanced micro-batches the per-mini-batch counts are
        # agreed over the ranks in ONE collective up front.
        minis = [s0 for _ in range(cfg.ppo_epochs) for s0 in range(0, n_total, mini)]
        counts_mb: List[Optional[int]] = [None] * len(minis)
        if cfg.use_dynamic_bsz:
            sums = [sum(lens[s0:s0 + mini]) if lens is not None else mini * t_len for s0 in minis]
            counts_mb = micro_batch_counts(sums, max(cfg.max_token_len_per_micro_batch, t_len), self.process_group,
                                           self.weight.device)
        plans: Dict[int, list] = {}

        def plan_of(i: int):
            if i not in plans:
                plans[i] = self._micro_plan(list(range(minis[i], minis[i] + mini)), lens, t_len, counts_mb[i])
            return plans[i]

        slot = 0
        if stager is not None and minis:
            stager.prefetch(0, batch, plan_of(0)[0][0])
        for i in range(len(minis)):
            plan = plan_of(i)
            for j, (rows, divisor, valid_rows) in enumerate(plan):
                if stager is not None:
                    nxt = plan[j + 1][0] if j + 1 < len(plan) else (plan_of(i + 1)[0][0] if i + 1 < len(minis) else None)
                    if nxt is not None:
                        stager.prefetch(slot ^ 1, batch, nxt)
                    micro = stager.get(slot)
                else:
                    micro = {k: _rows_of(v, rows) for k, v in batch.items()}
                micro.update({k: v[np.asarray(rows)] for k, v in non_tensor.items()})
                hidden = self._hidden(micro, train=True)
                step = grpo_micro_batch_step(
                    hidden.detach(), self.weight.detach(), micro["responses"], micro["old_log_probs"],
                    micro["advantages"], micro["ref_log_probs"] if use_ref else None, self._response_mask(micro),
                    temperature=temperature, clip_ratio_low=cfg.clip_ratio_low, clip_ratio_high=cfg.clip_ratio_high,
                    clip_ratio_dual=cfg.clip_ratio_dual, kl_penalty=cfg.kl_penalty if use_ref else None,
                    kl_coef=cfg.kl_coef, grad_accum=float(divisor), entropy_coeff=cfg.entropy_coeff,
                    want_entropy=cfg.log_true_entropy, dweight_accum=self.dweight,
                    valid_rows=valid_rows if self.compact_padding else None, defer=defer,
                )
                if hidden.requires_grad:
                    hidden.backward(step["dhidden"])  # continue into the transformer body
                else:
                    self.last_dhidden.append(step["dhidden"])
                if stager is not None:
                    stager.release(slot)
                    slot ^= 1
                pending.append(step["metrics"])
                step_of.append(len(norms))
                if j == 0 and i + 1 < len(minis):
                    plan_of(i + 1)  # host work while the GPU is busy with the micro-batch just enqueued
            if defer is not None:
                defer.flush()  # dW complete before it is all-reduced and applied
            norms.append(self._optimizer_step())
            plans.pop(i, None)

        # one device->host transfer for every scalar of this call
        host = torch.stack(pending).float().cpu() if pending else torch.zeros(0, _lib.NUM_METRICS)
        host_norms = torch.stack(norms).float().cpu().tolist() if norms else []
        metrics: Dict[str, Any] = defaultdict(list)
        rows = host.tolist()
        for i, row_vals in enumerate(rows):
            if use_ref:  # the reference ASSIGNS these two (dp_actor.py:274-275): last micro-batch wins
                metrics["actor/kl_loss"] = row_vals[_lib.MET_KL_LOSS]
                metrics["actor/kl_coef"] = cfg.kl_coef
            append_to_dict(metrics, {
                "actor/pg_loss": row_vals[_lib.MET_TOTAL],
                "actor/pg_clipfrac_higher": row_vals[_lib.MET_CLIPFRAC_HI],
                "actor/pg_clipfrac_lower": row_vals[_lib.MET_CLIPFRAC_LO],
                "actor/entropy_loss": row_vals[_lib.MET_ENTROPY],
                "actor/ppo_kl": row_vals[_lib.MET_PPO_KL],
            })
            if cfg.log_true_entropy:
                append_to_dict(metrics, {"actor/entropy": row_vals[_lib.MET_TRUE_ENTROPY]})
            if i + 1 == len(rows) or step_of[i + 1] != step_of[i]:
                append_to_dict(metrics, {"actor/grad_norm": host_norms[step_of[i]]})
        saturated = sum(r[_lib.MET_SATURATED] for r in rows)
        if saturated > 0:  # log p below -69: the label-referenced softmax of the fused head saturates there
            metrics["actor/saturated_tokens"] = saturated
            warnings.warn(f"{int(saturated)} unmasked tokens have log-probabilities below -69.3: their row sums may be "
                          "saturated in the fused head (a rollout / training mismatch?).")
        return metrics
