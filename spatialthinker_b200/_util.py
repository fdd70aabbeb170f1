"""Small host-side helpers shared by the reference-facing modules."""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import _lib

_scratch: Dict[Tuple[str, int], torch.Tensor] = {}


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """Every tensor must live on one CUDA device. There is deliberately no CPU path in the product."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "spatialthinker_b200 runs on CUDA tensors only (sm_100a kernels); there is no CPU fallback. "
                "The CPU restatement of this path lives in oracle/ and is test infrastructure."
            )
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError(f"tensors on different devices: {dev} and {t.device}")
    if dev is None:
        raise ValueError("no tensor arguments")
    return dev


def scratch(kind: str, device: torch.device, nbytes: int) -> torch.Tensor:
    """A cached, grow-only byte buffer per (kind, device): kernel workspaces are reused across calls.

    The buffer is shared by every call on that device, so calls must be ordered on ONE stream (the actor loop's: the
    library only enqueues on torch's current stream). Code that drives the head from several streams at once must hand
    each stream its own workspace through the C ABI (``grpo_*_workspace_bytes``) instead of this cache."""
    key = (kind, device.index if device.index is not None else torch.cuda.current_device())
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            del _scratch[key]
            del buf
        buf = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def release_scratch() -> None:
    _scratch.clear()


def f32c(t: torch.Tensor) -> torch.Tensor:
    """contiguous float32 view/copy (the reference's loss arithmetic is fp32: core_algos.py:408 `.float()`)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def mask_arg(mask: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """Masks are consumed in their own dtype (int64 attention-mask slices included) - no conversion pass."""
    if mask.dtype not in (torch.float32, torch.int64, torch.bool, torch.uint8):
        mask = mask.float()
    mask = mask.contiguous()
    return mask, _lib.mask_dtype_code(mask)
