"""Fused lm_head entry points: the logits tensor ``[tokens, vocab]`` is never materialised.

The reference reaches this arithmetic through three separate stages - HF ``lm_head`` (cuBLAS, logits written to HBM),
``logits.div_(temperature)`` and flash-attn's Triton cross-entropy (dp_actor.py:118-128, torch_functional.py:34-66) -
followed by ~25 elementwise kernels for the loss (core_algos.py:291-353, 394-436) and autograd's three backward
passes over the logits. A ``logits`` argument cannot be fused with the GEMM that produces it, so these functions take
the final hidden states and the ``lm_head`` weight instead:

* :func:`fused_lm_head_log_probs` - hidden, weight, labels -> log-probs (and entropy), differentiable.
* :func:`fused_grpo_loss` - one micro-batch of dp_actor.py:247-278: log-probs, clipped policy loss, KL term, masked
  means, and the gradients into hidden and weight, in a single pass of three GEMM units.
* :func:`grpo_micro_batch_step` - the same without autograd, accumulating ``dW`` straight into an fp32 buffer; this is
  what :mod:`spatialthinker_b200.dp_actor` loops over.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._util import f32c, mask_arg, require_cuda, scratch


def _check_head(hidden: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor):
    dev = require_cuda(hidden, weight, labels)
    if hidden.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise ValueError("hidden and weight must be bfloat16 (the actor's parameter dtype, actor/config.py:57)")
    if weight.dim() != 2 or hidden.shape[-1] != weight.shape[1]:
        raise ValueError(f"weight {tuple(weight.shape)} does not match hidden {tuple(hidden.shape)}")
    if labels.shape != hidden.shape[:-1]:
        raise ValueError(f"labels {tuple(labels.shape)} do not match hidden {tuple(hidden.shape)}")
    h2 = hidden.contiguous().view(-1, hidden.shape[-1])
    w2 = weight.contiguous()
    lab = labels.contiguous().view(-1).to(torch.int64)
    return dev, h2, w2, lab


# ----------------------------------------------------------------------------------------------------------------
# log-probs with autograd
# ----------------------------------------------------------------------------------------------------------------
class _FusedLogProbs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, weight, labels, temperature: float, want_entropy: bool):
        dev, h2, w2, lab = _check_head(hidden, weight, labels)
        lib = _lib.load()
        rows, hdim = h2.shape
        vocab = w2.shape[0]
        logp = torch.empty(rows, dtype=torch.float32, device=dev)
        ent = torch.empty(rows, dtype=torch.float32, device=dev) if want_entropy else None
        nbytes = lib.grpo_lmhead_fwd_workspace_bytes(rows, hdim, vocab)
        ws = scratch("head", dev, nbytes)
        with torch.cuda.device(dev):
            _lib.check(
                lib.grpo_lmhead_logprob_fwd(h2.data_ptr(), w2.data_ptr(), lab.data_ptr(), rows, hdim, vocab,
                                            float(temperature), logp.data_ptr(), _lib.ptr(ent), None, ws.data_ptr(),
                                            ws.numel(), _lib.stream_ptr(dev)),
                "grpo_lmhead_logprob_fwd",
            )
        ctx.save_for_backward(h2, w2, lab)
        ctx.temperature = float(temperature)
        ctx.hshape, ctx.wdtype = hidden.shape, weight.dtype
        ctx.want_entropy = want_entropy
        lead = hidden.shape[:-1]
        return logp.view(*lead), (ent.view(*lead) if want_entropy else None)

    @staticmethod
    def backward(ctx, g_logp, g_ent):
        h2, w2, lab = ctx.saved_tensors
        lib = _lib.load()
        dev = h2.device
        rows, hdim = h2.shape
        vocab = w2.shape[0]
        gl = f32c(g_logp.reshape(-1)) if g_logp is not None else torch.zeros(rows, dtype=torch.float32, device=dev)
        ge = f32c(g_ent.reshape(-1)) if (g_ent is not None and ctx.want_entropy) else None
        dh = torch.empty(rows, hdim, dtype=torch.bfloat16, device=dev)
        dw = torch.zeros(vocab, hdim, dtype=torch.float32, device=dev)
        nbytes = lib.grpo_lmhead_bwd_workspace_bytes(rows, hdim, vocab)
        ws = scratch("head", dev, nbytes)
        with torch.cuda.device(dev):
            _lib.check(
                lib.grpo_lmhead_bwd(h2.data_ptr(), w2.data_ptr(), lab.data_ptr(), gl.data_ptr(), _lib.ptr(ge), rows,
                                    hdim, vocab, ctx.temperature, dh.data_ptr(), dw.data_ptr(), ws.data_ptr(),
                                    ws.numel(), _lib.stream_ptr(dev)),
                "grpo_lmhead_bwd",
            )
        return dh.view(ctx.hshape), dw.to(ctx.wdtype), None, None, None


def fused_lm_head_log_probs(
    hidden: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor, temperature: float = 1.0, want_entropy: bool = False
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """``log_softmax(hidden @ weight.T / temperature)[labels]`` (and the entropy of that softmax) without the logits.

    hidden ``[..., H]`` bf16, weight ``[V, H]`` bf16, labels ``[...]`` int64 -> (log-probs ``[...]`` fp32,
    entropy ``[...]`` fp32 or None). Differentiable with respect to ``hidden`` and ``weight``; the backward recomputes
    the logits tiles chunk by chunk (4 GEMM units in total). Forward-only use (``compute_log_prob``,
    dp_actor.py:170-210) is a single GEMM unit.
    """
    return _FusedLogProbs.apply(hidden, weight, labels, temperature, want_entropy)


# ----------------------------------------------------------------------------------------------------------------
# deferred dW for small micro-batches
# ----------------------------------------------------------------------------------------------------------------
class DeferredDW:
    """One chunk workspace shared by several SMALL micro-batches of a gradient-accumulation loop (dp_actor.py:242-290).

    ``dW`` of the loop is a sum over its micro-batches, so the GEMM that produces it need not run once per micro-batch
    (the reference ships ``micro_batch_size_per_device_for_update: 4``, scripts/config.yaml:28). Each micro-batch handed
    to :func:`grpo_micro_batch_step` with ``defer=session`` runs forward + loss + dHidden in its own 512-row-aligned
    window of the session's exp-stash; :meth:`flush` runs ONE dW GEMM over everything collected - ``dW`` is
    read-modify-written once and the GEMM's K dimension is long enough to hide its fp32 drain. Everything but the
    stash-dependent part of ``dW`` (log-probs, metrics, dHidden) is final when the step returns. Call :meth:`flush`
    before ``dweight`` is read (optimizer step, all-reduce); a micro-batch that does not fit flushes first, one larger
    than the workspace takes the ordinary path. No entropy gradient on this path (``entropy_coeff`` must be 0).
    """

    def __init__(self, weight: torch.Tensor, dweight_accum: torch.Tensor):
        dev = require_cuda(weight, dweight_accum)
        if dweight_accum.shape != weight.shape or dweight_accum.dtype != torch.float32 or not dweight_accum.is_contiguous():
            raise ValueError("dweight_accum must be a contiguous float32 tensor shaped like weight")
        lib = _lib.load()
        self.device = dev
        self.dweight = dweight_accum
        self.vocab, self.hdim = weight.shape
        # whole 512-row tiles only (the chunk_rows knob may have been set to anything); 0 disables deferral
        self.capacity = int(lib.grpo_chunk_capacity_rows()) // 512 * 512
        self.nbytes = int(lib.grpo_fused_loss_workspace_bytes(self.capacity, self.hdim, self.vocab)) if self.capacity else 0
        # its own buffer: the stash must survive other head calls (compute_log_prob ...) between two micro-batches.
        # Allocated by the first reserve() that succeeds - an actor whose micro-batches all exceed the capacity never
        # pays for it.
        self.workspace: Optional[torch.Tensor] = None
        self.next_row0 = 0   # where the next slot starts (multiple of 512)
        self.total_rows = 0  # end of the last slot
        self.pending = 0     # micro-batches collected since the last flush

    def reserve(self, rows: int) -> Optional[int]:
        """Start row of a slot for ``rows`` rows (flushing first when the workspace is full); None when ``rows`` can
        never fit and the caller should take the ordinary path."""
        if rows <= 0 or rows > self.capacity:
            return None
        if self.workspace is None:
            # zeroed once: stash rows no launch has written yet meet zero operand rows in the dW GEMM and must be finite
            self.workspace = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
        if self.next_row0 + rows > self.capacity:
            self.flush()
        row0 = self.next_row0
        self.total_rows = row0 + rows
        self.next_row0 = -(-self.total_rows // 512) * 512
        self.pending += 1
        return row0

    def flush(self) -> None:
        """dweight += stash^T . scaled hidden over every row collected since the last flush."""
        if self.pending:
            with torch.cuda.device(self.device):
                _lib.check(
                    _lib.load().grpo_deferred_dw_flush(self.total_rows, self.capacity, self.hdim, self.vocab,
                                                       self.dweight.data_ptr(), self.workspace.data_ptr(),
                                                       self.workspace.numel(), _lib.stream_ptr(self.device)),
                    "grpo_deferred_dw_flush",
                )
        self.next_row0 = self.total_rows = self.pending = 0

    def release(self) -> None:
        """Flush what is pending and give the chunk workspace back (it is re-allocated by the next small micro-batch)."""
        self.flush()
        self.workspace = None


# ----------------------------------------------------------------------------------------------------------------
# one micro-batch of the actor update, no autograd
# ----------------------------------------------------------------------------------------------------------------
METRIC_KEYS = {
    "actor/pg_loss": _lib.MET_TOTAL,  # reported after the KL term was added, dp_actor.py:271,281
    "actor/pg_clipfrac_higher": _lib.MET_CLIPFRAC_HI,
    "actor/pg_clipfrac_lower": _lib.MET_CLIPFRAC_LO,
    "actor/entropy_loss": _lib.MET_ENTROPY,
    "actor/ppo_kl": _lib.MET_PPO_KL,
}


def grpo_micro_batch_step(
    hidden: torch.Tensor,
    weight: torch.Tensor,
    labels: torch.Tensor,
    old_log_probs: torch.Tensor,
    advantages: torch.Tensor,
    ref_log_probs: Optional[torch.Tensor],
    response_mask: torch.Tensor,
    *,
    temperature: float = 1.0,
    clip_ratio_low: float = 0.2,
    clip_ratio_high: float = 0.3,
    clip_ratio_dual: float = 3.0,
    kl_penalty: Optional[str] = "low_var_kl",
    kl_coef: float = 0.0,
    grad_accum: float = 1.0,
    entropy_coeff: float = 0.0,
    want_entropy: bool = False,
    dweight_accum: Optional[torch.Tensor] = None,
    need_grads: bool = True,
    valid_rows: Optional[int] = None,
    defer: Optional[DeferredDW] = None,
) -> Dict[str, torch.Tensor]:
    """Forward + backward of one micro-batch (dp_actor.py:247-278) entirely on the device, no host sync.

    Returns a dict of device tensors: ``log_probs`` [.. ] fp32, ``entropy`` (or None), ``metrics`` (fp32 vector indexed
    by the ``_lib.MET_*`` slots), ``dhidden`` (bf16, gradient of ``loss = total / grad_accum``) and ``dweight`` - the
    fp32 ``[V, H]`` buffer the weight gradient was ACCUMULATED into (``dweight_accum`` if given, else a fresh zero
    buffer).

    ``valid_rows`` (optional, a HOST integer = number of non-zero mask entries, e.g. the sum of the response lengths the
    trainer already knows): padded positions are then dropped before the GEMMs - the reference computes and discards
    them (dp_actor.py:136-139) - and the outputs are scattered back (``log_probs`` / ``entropy`` are 0 and ``dhidden``
    rows are 0 at padded positions). Without the hint nothing is compacted, because finding the count would cost a
    device->host sync.

    ``defer`` (optional :class:`DeferredDW` built on ``dweight_accum``): the stash-dependent part of ``dweight`` is left
    to ``defer.flush()``; everything else is final on return.
    """
    if valid_rows is not None and 0 < valid_rows < response_mask.numel():
        return _compacted_step(hidden, weight, labels, old_log_probs, advantages, ref_log_probs, response_mask,
                               int(valid_rows), dict(temperature=temperature, clip_ratio_low=clip_ratio_low,
                                                     clip_ratio_high=clip_ratio_high, clip_ratio_dual=clip_ratio_dual,
                                                     kl_penalty=kl_penalty, kl_coef=kl_coef, grad_accum=grad_accum,
                                                     entropy_coeff=entropy_coeff, want_entropy=want_entropy,
                                                     dweight_accum=dweight_accum, need_grads=need_grads, defer=defer))
    dev, h2, w2, lab = _check_head(hidden, weight, labels)
    lib = _lib.load()
    rows, hdim = h2.shape
    vocab = w2.shape[0]
    lead = hidden.shape[:-1]
    for name, t in (("old_log_probs", old_log_probs), ("advantages", advantages), ("response_mask", response_mask)):
        if t.shape != lead:
            raise ValueError(f"{name} {tuple(t.shape)} does not match the token layout {tuple(lead)}")
    use_kl = ref_log_probs is not None and kl_penalty is not None
    mode = _lib.KL_MODES.get(kl_penalty, None) if use_kl else -1
    if mode is None:
        raise NotImplementedError(f"Unknown KL penalty: {kl_penalty}.")
    old, adv = f32c(old_log_probs).view(-1), f32c(advantages).view(-1)
    ref = f32c(ref_log_probs).view(-1) if use_kl else None
    mask, code = mask_arg(response_mask)
    want_entropy = want_entropy or entropy_coeff != 0.0
    logp = torch.empty(rows, dtype=torch.float32, device=dev)
    ent = torch.empty(rows, dtype=torch.float32, device=dev) if want_entropy else None
    metrics = torch.empty(_lib.NUM_METRICS, dtype=torch.float32, device=dev)
    dh = dw = None
    if need_grads:
        dh = torch.empty(rows, hdim, dtype=torch.bfloat16, device=dev)
        if dweight_accum is not None:
            if dweight_accum.shape != w2.shape or dweight_accum.dtype != torch.float32 or not dweight_accum.is_contiguous():
                raise ValueError("dweight_accum must be a contiguous float32 tensor shaped like weight")
            dw = dweight_accum
        else:
            dw = torch.zeros(vocab, hdim, dtype=torch.float32, device=dev)
    slot_row0 = None
    if defer is not None:
        if not need_grads or dw is not defer.dweight or entropy_coeff != 0.0:
            raise ValueError("defer needs need_grads=True, dweight_accum=defer.dweight and entropy_coeff == 0")
        if (defer.hdim, defer.vocab) != (hdim, vocab) or defer.device != dev:
            raise ValueError("defer was built for another weight")
        before = (defer.next_row0, defer.total_rows, defer.pending)
        slot_row0 = defer.reserve(rows)  # None: larger than the workspace - ordinary path below
    if slot_row0 is not None:
        with torch.cuda.device(dev):
            rc = lib.grpo_fused_loss_fwd_bwd_slot(
                h2.data_ptr(), w2.data_ptr(), lab.data_ptr(), old.data_ptr(), adv.data_ptr(), _lib.ptr(ref),
                mask.data_ptr(), code, rows, hdim, vocab, float(temperature), float(clip_ratio_low),
                float(clip_ratio_high), float(clip_ratio_dual), mode, float(kl_coef if use_kl else 0.0),
                float(grad_accum), logp.data_ptr(), _lib.ptr(ent), dh.data_ptr(), dw.data_ptr(),
                metrics.data_ptr(), slot_row0, defer.capacity, defer.workspace.data_ptr(),
                defer.workspace.numel(), _lib.stream_ptr(dev))
        if rc != 0:
            # the slot was not (completely) filled: give it back, or the flush would sum whatever an earlier group left
            # there. (slot_row0 != before[0]: reserve() flushed first and this was the first slot of a new group.)
            defer.next_row0, defer.total_rows, defer.pending = before if slot_row0 == before[0] else (0, 0, 0)
        _lib.check(rc, "grpo_fused_loss_fwd_bwd_slot")
        return {
            "log_probs": logp.view(*lead),
            "entropy": ent.view(*lead) if ent is not None else None,
            "metrics": metrics,
            "dhidden": dh.view(hidden.shape),
            "dweight": dw,
            "used_kl": use_kl,
        }
    nbytes = lib.grpo_fused_loss_workspace_bytes(rows, hdim, vocab)
    ws = scratch("head", dev, nbytes)
    with torch.cuda.device(dev):
        _lib.check(
            lib.grpo_fused_loss_fwd_bwd(
                h2.data_ptr(), w2.data_ptr(), lab.data_ptr(), old.data_ptr(), adv.data_ptr(), _lib.ptr(ref),
                mask.data_ptr(), code, rows, hdim, vocab, float(temperature), float(clip_ratio_low),
                float(clip_ratio_high), float(clip_ratio_dual), mode, float(kl_coef if use_kl else 0.0),
                float(entropy_coeff), float(grad_accum), logp.data_ptr(), _lib.ptr(ent), _lib.ptr(dh), _lib.ptr(dw),
                metrics.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)),
            "grpo_fused_loss_fwd_bwd",
        )
    return {
        "log_probs": logp.view(*lead),
        "entropy": ent.view(*lead) if ent is not None else None,
        "metrics": metrics,
        "dhidden": dh.view(hidden.shape) if dh is not None else None,
        "dweight": dw,
        "used_kl": use_kl,
    }


def compact_index(response_mask: torch.Tensor):
    """(gather_idx, inverse, count) of the unmasked slots of ``response_mask`` (flattened), built on the device by
    ``grpo_compact_index`` - stable order, no host synchronisation. ``count`` is a 1-element int32 device tensor."""
    dev = require_cuda(response_mask)
    lib = _lib.load()
    mask, code = mask_arg(response_mask)
    n = mask.numel()
    gather_idx = torch.empty(n, dtype=torch.int32, device=dev)
    inverse = torch.empty(n, dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    nbytes = lib.grpo_compact_scratch_bytes(n)
    tmp = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.grpo_compact_index(mask.data_ptr(), code, n, gather_idx.data_ptr(), inverse.data_ptr(),
                                          count.data_ptr(), tmp.data_ptr(), nbytes, _lib.stream_ptr(dev)),
                   "grpo_compact_index")
    return gather_idx, inverse, count


def gather_rows(src: torch.Tensor, gather_idx: torch.Tensor, m: int) -> torch.Tensor:
    """``out[j] = src[gather_idx[j]]`` for ``j < m`` over the leading dimension (``grpo_gather_rows``)."""
    dev = require_cuda(src, gather_idx)
    src = src.contiguous()
    row_bytes = math.prod(src.shape[1:]) * src.element_size()
    out = torch.empty((m,) + tuple(src.shape[1:]), dtype=src.dtype, device=dev)
    if m == 0:
        return out
    with torch.cuda.device(dev):
        _lib.check(_lib.load().grpo_gather_rows(src.data_ptr(), gather_idx.data_ptr(), m, row_bytes, out.data_ptr(),
                                                _lib.stream_ptr(dev)), "grpo_gather_rows")
    return out


def scatter_rows(src: torch.Tensor, inverse: torch.Tensor) -> torch.Tensor:
    """``out[i] = src[inverse[i]]`` where ``inverse[i] >= 0``, zeros elsewhere (``grpo_scatter_rows``)."""
    dev = require_cuda(src, inverse)
    src = src.contiguous()
    n = inverse.numel()
    row_bytes = math.prod(src.shape[1:]) * src.element_size()
    if src.shape[0] == 0 or n == 0:  # nothing was kept: all zeros
        return torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=dev)
    out = torch.empty((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().grpo_scatter_rows(src.data_ptr(), inverse.data_ptr(), n, row_bytes, out.data_ptr(),
                                                 _lib.stream_ptr(dev)), "grpo_scatter_rows")
    return out


def _compacted_step(hidden, weight, labels, old_log_probs, advantages, ref_log_probs, response_mask, valid_rows, kw):
    """Gather the ``valid_rows`` unmasked token rows, run the fused step on them, scatter the results back (padded slots
    get zeros). Index, gather and scatter are this library's kernels (csrc/compact_kernels.cuh); the count comes from the
    host, so nothing synchronises."""
    lead = hidden.shape[:-1]
    hdim = hidden.shape[-1]
    gather_idx, inverse, _ = compact_index(response_mask)
    take = lambda t: None if t is None else gather_rows(t.reshape(-1), gather_idx, valid_rows)  # noqa: E731
    mask_flat = response_mask.reshape(-1)
    if mask_flat.element_size() % 4 != 0:  # bool / uint8 masks: the row kernels move 4-byte words
        mask_flat = mask_flat.to(torch.float32)
    res = grpo_micro_batch_step(
        gather_rows(hidden.reshape(-1, hdim), gather_idx, valid_rows), weight, take(labels), take(f32c(old_log_probs)),
        take(f32c(advantages)), take(None if ref_log_probs is None else f32c(ref_log_probs)), take(mask_flat), **kw)
    res["log_probs"] = scatter_rows(res["log_probs"], inverse).view(*lead)
    if res["entropy"] is not None:
        res["entropy"] = scatter_rows(res["entropy"], inverse).view(*lead)
    if res["dhidden"] is not None:
        res["dhidden"] = scatter_rows(res["dhidden"], inverse).view(hidden.shape)
    return res


def metrics_to_dict(metrics: torch.Tensor, used_kl: bool, kl_coef: float) -> Dict[str, float]:
    """ONE device->host read for all of a micro-batch's ``actor/*`` scalars (the reference does 5-6 ``.item()`` syncs
    per micro-batch, dp_actor.py:274-286)."""
    host = metrics.detach().float().cpu().tolist()
    out = {key: host[slot] for key, slot in METRIC_KEYS.items()}
    if used_kl:
        out["actor/kl_loss"] = host[_lib.MET_KL_LOSS]
        out["actor/kl_coef"] = kl_coef
    return out


# ----------------------------------------------------------------------------------------------------------------
# the same with autograd
# ----------------------------------------------------------------------------------------------------------------
class _FusedGrpoLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, weight, labels, old_log_probs, advantages, ref_log_probs, response_mask, kw):
        need = hidden.requires_grad or weight.requires_grad
        res = grpo_micro_batch_step(hidden, weight, labels, old_log_probs, advantages, ref_log_probs, response_mask,
                                    need_grads=need, **kw)
        if need:
            ctx.save_for_backward(res["dhidden"], res["dweight"])
        ctx.wdtype = weight.dtype
        ctx.need = need
        m = res["metrics"]
        ctx.mark_non_differentiable(m, res["log_probs"])
        ent = res["entropy"]
        if ent is not None:
            ctx.mark_non_differentiable(ent)
        ctx.used_kl = res["used_kl"]
        return m[_lib.MET_SCALED].clone(), m, res["log_probs"], ent

    @staticmethod
    def backward(ctx, g_loss, g_m, g_lp, g_ent):
        if not ctx.need:
            return (None,) * 8
        dh, dw = ctx.saved_tensors
        # gradients were produced in the forward pass for d(loss) = 1; apply the upstream scalar here
        return (dh.float() * g_loss).to(dh.dtype), (dw * g_loss).to(ctx.wdtype), None, None, None, None, None, None


def fused_grpo_loss(
    hidden: torch.Tensor,
    weight: torch.Tensor,
    labels: torch.Tensor,
    old_log_probs: torch.Tensor,
    advantages: torch.Tensor,
    ref_log_probs: Optional[torch.Tensor],
    response_mask: torch.Tensor,
    *,
    temperature: float = 1.0,
    clip_ratio_low: float = 0.2,
    clip_ratio_high: float = 0.3,
    clip_ratio_dual: float = 3.0,
    kl_penalty: Optional[str] = "low_var_kl",
    kl_coef: float = 0.0,
    grad_accum: float = 1.0,
    entropy_coeff: float = 0.0,
    want_entropy: bool = False,
) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """One micro-batch of the actor update as a differentiable scalar.

    ``loss = (pg_loss + kl_coef * masked_mean(kl) - entropy_coeff * masked_mean(entropy)) / grad_accum`` with the
    reference's per-micro-batch masked means (dp_actor.py:253-277). ``loss.backward()`` delivers the gradients that
    were already produced in the forward pass (three GEMM units in total, nothing recomputed).

    Returns ``(loss, metrics)``; ``metrics`` holds 0-d device tensors under the reference's ``actor/*`` keys
    (dp_actor.py:274-286) plus ``log_probs`` (and ``entropy`` when requested). Nothing is copied to the host.
    """
    kw = dict(temperature=temperature, clip_ratio_low=clip_ratio_low, clip_ratio_high=clip_ratio_high,
              clip_ratio_dual=clip_ratio_dual, kl_penalty=kl_penalty, kl_coef=kl_coef, grad_accum=grad_accum,
              entropy_coeff=entropy_coeff, want_entropy=want_entropy)
    loss, m, logp, ent = _FusedGrpoLoss.apply(hidden, weight, labels, old_log_probs, advantages, ref_log_probs,
                                              response_mask, kw)
    metrics = {key: m[slot] for key, slot in METRIC_KEYS.items()}
    if ref_log_probs is not None and kl_penalty is not None:
        metrics["actor/kl_loss"] = m[_lib.MET_KL_LOSS]
        metrics["actor/kl_coef"] = kl_coef
    metrics["log_probs"] = logp
    if ent is not None:
        metrics["entropy"] = ent
        metrics["actor/entropy"] = m[_lib.MET_TRUE_ENTROPY]
    return loss, metrics
