"""Drop-in for the hot-path part of ``verl/utils/torch_functional.py`` (lines 26-97 of the reference).

Same names, argument meaning and return conventions as the reference; the arithmetic runs in hand-written CUDA
(``csrc/logits_kernels.cuh``, ``csrc/loss_kernels.cuh``) through the C ABI. The upstream-veRL spellings named by the
task (``logprobs_from_logits``, ``entropy_from_logits``) are exported as well.

These functions take MATERIALISED logits, as the reference does. The path that never builds the logits tensor is
``spatialthinker_b200.fused`` - a ``logits`` argument cannot be fused with the GEMM that produces it.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._util import f32c, mask_arg, require_cuda

# The reference calls flash-attn's cross entropy with inplace_backward=True (torch_functional.py:36): the gradient
# overwrites the logits buffer. Off by default here (safe if logits are reused); set True for the reference's memory
# behaviour.
INPLACE_BACKWARD = False

_LOGITS_DTYPES = {torch.float32: _lib.LOGITS_F32, torch.bfloat16: _lib.LOGITS_BF16, torch.float16: _lib.LOGITS_F16}


def _as_rows(logits: torch.Tensor):
    if logits.dtype not in _LOGITS_DTYPES:
        raise ValueError(f"unsupported logits dtype {logits.dtype}")
    vocab = logits.shape[-1]
    z = logits.contiguous().view(-1, vocab)  # same flattening as torch_functional.py:57-60
    return z, vocab


class _LogProbsFromLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: torch.Tensor, labels: torch.Tensor, want_entropy: bool):
        dev = require_cuda(logits, labels)
        lib = _lib.load()
        z, vocab = _as_rows(logits)
        lab = labels.contiguous().view(-1).to(torch.int64)
        rows = z.shape[0]
        if lab.numel() != rows:
            raise ValueError(f"labels {tuple(labels.shape)} do not match logits {tuple(logits.shape)}")
        logp = torch.empty(rows, dtype=torch.float32, device=dev)
        lse = torch.empty(rows, dtype=torch.float32, device=dev)
        ent = torch.empty(rows, dtype=torch.float32, device=dev) if want_entropy else None
        with torch.cuda.device(dev):
            _lib.check(
                lib.grpo_logprob_from_logits(z.data_ptr(), _LOGITS_DTYPES[z.dtype], lab.data_ptr(), rows, vocab,
                                             z.stride(0), logp.data_ptr(), _lib.ptr(ent), lse.data_ptr(),
                                             _lib.stream_ptr(dev)),
                "grpo_logprob_from_logits",
            )
        ctx.save_for_backward(z, lab, lse, ent)
        ctx.shape = logits.shape
        lead = logits.shape[:-1]
        if want_entropy:
            return logp.view(*lead), ent.view(*lead)
        return logp.view(*lead), None

    @staticmethod
    def backward(ctx, g_logp, g_ent):
        z, lab, lse, ent = ctx.saved_tensors
        lib = _lib.load()
        dev = z.device
        rows, vocab = z.shape
        out = z if INPLACE_BACKWARD else torch.empty_like(z)
        gl = f32c(g_logp.reshape(-1)) if g_logp is not None else None
        ge = f32c(g_ent.reshape(-1)) if (g_ent is not None and ent is not None) else None
        with torch.cuda.device(dev):
            _lib.check(
                lib.grpo_logprob_from_logits_bwd(z.data_ptr(), _LOGITS_DTYPES[z.dtype], lab.data_ptr(), lse.data_ptr(),
                                                 _lib.ptr(gl), _lib.ptr(ge), _lib.ptr(ent), rows, vocab, z.stride(0),
                                                 out.data_ptr(), out.stride(0), _lib.stream_ptr(dev)),
                "grpo_logprob_from_logits_bwd",
            )
        return out.view(ctx.shape), None, None


def log_probs_from_logits(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """Log-prob of ``labels`` under ``logits``: fp32, shape ``logits.shape[:-1]``, negative numbers.

    Reference: verl/utils/torch_functional.py:45-66 (training branch :34-42: ``-cross_entropy_loss(...)``).
    """
    return _LogProbsFromLogits.apply(logits, labels, False)[0]


def entropy_from_logits(logits: torch.Tensor) -> torch.Tensor:
    """Per-token entropy ``logsumexp(z) - sum softmax(z) * z`` (upstream-veRL name; the reference only logs the
    estimator ``-masked_mean(log_probs)``, dp_actor.py:253)."""
    dummy = torch.zeros(logits.shape[:-1], dtype=torch.int64, device=logits.device)
    return _LogProbsFromLogits.apply(logits, dummy, True)[1]


logprobs_from_logits = log_probs_from_logits  # upstream-veRL spelling used by BASELINE.json's north_star


class _MaskedMeanAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, values: torch.Tensor, mask: torch.Tensor, eps: float):
        dev = require_cuda(values, mask)
        lib = _lib.load()
        x = f32c(values).view(-1)
        m, code = mask_arg(mask.expand_as(values) if mask.shape != values.shape else mask)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        acc = torch.empty(2, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.grpo_masked_mean(x.data_ptr(), m.data_ptr(), code, x.numel(), float(eps), out.data_ptr(),
                                            acc.data_ptr(), _lib.stream_ptr(dev)), "grpo_masked_mean")
        ctx.save_for_backward(m, acc)
        ctx.eps = eps
        ctx.vshape, ctx.vdtype = values.shape, values.dtype
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        m, acc = ctx.saved_tensors
        denom = acc[1].float() + ctx.eps
        return (g * m.view(ctx.vshape).float() / denom).to(ctx.vdtype), None, None


def masked_mean(values: torch.Tensor, mask: torch.Tensor, dim: Optional[int] = None, eps: float = 1e-8) -> torch.Tensor:
    """``sum(values * mask) / (sum(mask) + eps)`` - verl/utils/torch_functional.py:69-71.

    ``dim=None`` (every use on the hot path, dp_actor.py:253-270, core_algos.py:349-352) runs the reduction kernel;
    a per-dimension mean (only ``apply_kl_penalty``, ray_trainer.py:141) is the same expression in torch ops.
    """
    if dim is None:
        return _MaskedMeanAll.apply(values, mask, eps)
    require_cuda(values, mask)
    return (values * mask).sum(dim=dim) / (mask.sum(dim=dim) + eps)


def _masked_moments(values: torch.Tensor, mask: torch.Tensor, unbiased: bool) -> torch.Tensor:
    dev = require_cuda(values, mask)
    lib = _lib.load()
    x = f32c(values).view(-1)
    m, code = mask_arg(mask.expand_as(values) if mask.shape != values.shape else mask)
    out = torch.empty(2, dtype=torch.float32, device=dev)
    acc = torch.empty(4, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.grpo_masked_var(x.data_ptr(), m.data_ptr(), code, x.numel(), int(bool(unbiased)), out.data_ptr(),
                                       acc.data_ptr(), _lib.stream_ptr(dev)), "grpo_masked_var")
    return out


@torch.no_grad()
def masked_var(values: torch.Tensor, mask: torch.Tensor, unbiased: bool = True) -> torch.Tensor:
    """Variance over the masked entries - verl/utils/torch_functional.py:74-89 (Bessel's correction unless
    ``sum(mask) <= 1``, where the reference prints a warning and returns the biased value). Not differentiable here:
    the reference only uses it under ``torch.no_grad`` (advantage whitening)."""
    return _masked_moments(values, mask, unbiased)[0]


@torch.no_grad()
def masked_whiten(values: torch.Tensor, mask: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """``(values - masked_mean) * rsqrt(masked_var + eps)`` at every position - torch_functional.py:92-95."""
    dev = require_cuda(values, mask)
    lib = _lib.load()
    x = f32c(values)
    m, code = mask_arg(mask.expand_as(values) if mask.shape != values.shape else mask)
    out = torch.empty_like(x)
    acc = torch.empty(4, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.grpo_masked_whiten(x.data_ptr(), m.data_ptr(), code, x.numel(), float(eps), out.data_ptr(),
                                          acc.data_ptr(), _lib.stream_ptr(dev)), "grpo_masked_whiten")
    return out
