"""ctypes binding of ``libgrpo_b200.so`` (C ABI: ``include/grpo_b200.h``).

This is the only place the shared library is opened. There is no fallback: if the library is missing, or an entry
point returns non-zero, the caller gets an exception - the product path never routes around the CUDA kernels.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# GRPO_B200_LIB: measurement tooling only (e.g. the -DGRPO_TRACE build of tools/trace_tiles.py); the product loads the
# in-tree library
LIB_PATH = os.environ.get("GRPO_B200_LIB") or os.path.join(_HERE, "libgrpo_b200.so")

# keep in sync with include/grpo_b200.h
KL_MODES = {None: -1, "none": -1, "low_var_kl": 0, "kl": 1, "abs": 2, "mse": 3, "chi2": 4}
MASK_F32, MASK_I64, MASK_U8, MASK_NONE = 0, 1, 2, 3
LOGITS_F32, LOGITS_BF16, LOGITS_F16 = 0, 1, 2
NUM_METRICS = 11
NUM_PHASES = 6
PHASE_NAMES = ("logits_gemm", "row_stats", "token_loss", "grad_prep", "dhidden_gemm", "dweight_gemm")
MET_PG_LOSS, MET_CLIPFRAC_HI, MET_CLIPFRAC_LO, MET_PPO_KL, MET_KL_LOSS = 0, 1, 2, 3, 4
MET_ENTROPY, MET_TOTAL, MET_SCALED, MET_TRUE_ENTROPY, MET_MASK_SUM, MET_SATURATED = 5, 6, 7, 8, 9, 10
GRAD_SCRATCH_DOUBLES = 1024
ABI_VERSION = 2

_P = c_void_p
_SIGNATURES = {
    "grpo_abi_version": (c_int, []),
    "grpo_last_error": (c_char_p, []),
    "grpo_launch_count": (ctypes.c_longlong, []),
    "grpo_set_option": (c_int, [c_char_p, c_int]),
    "grpo_debug_probe_offset": (c_size_t, [c_int64, c_int64, c_int64, c_int]),
    "grpo_profile_enable": (c_int, [c_int]),
    "grpo_profile_read": (c_int, [_P, _P, c_int]),
    "grpo_lmhead_fwd_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "grpo_lmhead_logprob_fwd": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_float, _P, _P, _P, _P, c_size_t, _P]),
    "grpo_lmhead_bwd_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "grpo_lmhead_bwd": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int64, c_float, _P, _P, _P, c_size_t, _P]),
    "grpo_fused_loss_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "grpo_fused_loss_fwd_bwd": (
        c_int,
        [_P, _P, _P, _P, _P, _P, _P, c_int, c_int64, c_int64, c_int64, c_float, c_float, c_float, c_float, c_int,
         c_float, c_float, c_float, _P, _P, _P, _P, _P, _P, c_size_t, _P],
    ),
    "grpo_chunk_capacity_rows": (ctypes.c_longlong, []),
    "grpo_fused_loss_fwd_bwd_slot": (
        c_int,
        [_P, _P, _P, _P, _P, _P, _P, c_int, c_int64, c_int64, c_int64, c_float, c_float, c_float, c_float, c_int,
         c_float, c_float, _P, _P, _P, _P, _P, c_int64, c_int64, _P, c_size_t, _P],
    ),
    "grpo_deferred_dw_flush": (c_int, [c_int64, c_int64, c_int64, c_int64, _P, _P, c_size_t, _P]),
    "grpo_grad_sumsq": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P]),
    "grpo_grad_scale_cast": (c_int, [_P, c_int64, _P, c_float, _P, c_int, _P]),
    "grpo_ipc_export": (c_int, [_P, _P, _P]),
    "grpo_ipc_open": (c_int, [_P, c_int64, _P]),
    "grpo_ipc_close": (c_int, [_P, c_int64]),
    "grpo_peer_barrier": (c_int, [_P, c_int, c_int, ctypes.c_uint, c_int, _P]),
    "grpo_peer_reduce_scatter_sumsq": (c_int, [_P, _P, c_int, c_int, c_int64, _P, _P]),
    "grpo_peer_scale_cast_allgather": (c_int, [_P, _P, c_int, c_int, c_int64, _P, c_float, c_int, _P]),
    "grpo_peer_allreduce_mean": (c_int, [_P, c_int, c_int, c_int64, _P]),
    "grpo_debug_peer_slab": (c_int, [c_int64, c_int, c_int, _P, _P]),
    "grpo_policy_loss_fwd_bwd": (
        c_int,
        [_P, _P, _P, _P, _P, c_int, c_int64, c_float, c_float, c_float, c_int, c_float, c_float, _P, _P, _P, _P],
    ),
    "grpo_compute_kl": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P]),
    "grpo_masked_mean": (c_int, [_P, _P, c_int, c_int64, c_float, _P, _P, _P]),
    "grpo_advantage": (c_int, [_P, _P, c_int, _P, _P, c_int64, c_int64, c_int64, c_float, _P, _P, _P]),
    "grpo_sequence_scores": (c_int, [_P, c_int64, c_int64, _P, _P]),
    "grpo_advantage_from_scores": (c_int, [_P, _P, _P, c_int64, c_int64, c_float, c_int64, _P, c_int, c_int64, c_int64, _P, _P, _P]),
    "grpo_rloo_advantage": (c_int, [_P, _P, c_int, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P]),
    "grpo_remax_advantage": (c_int, [_P, _P, _P, c_int, c_int64, c_int64, _P, _P, _P]),
    "grpo_reinforce_pp_advantage": (c_int, [_P, _P, c_int, c_int64, c_int64, c_float, _P, _P, _P, _P]),
    "grpo_gae_advantage": (c_int, [_P, _P, _P, c_int, c_int64, c_int64, c_float, c_float, _P, _P, _P, _P]),
    "grpo_masked_whiten": (c_int, [_P, _P, c_int, c_int64, c_float, _P, _P, _P]),
    "grpo_masked_var": (c_int, [_P, _P, c_int, c_int64, c_int, _P, _P, _P]),
    "grpo_value_loss_fwd_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int64, c_float, _P, _P, _P, _P]),
    "grpo_kl_penalty_rewards": (c_int, [_P, _P, _P, _P, c_int, c_int64, c_int64, c_int, c_float, _P, _P, _P, _P]),
    "grpo_compact_scratch_bytes": (c_size_t, [c_int64]),
    "grpo_compact_index": (c_int, [_P, c_int, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "grpo_gather_rows": (c_int, [_P, _P, c_int64, c_int64, _P, _P]),
    "grpo_scatter_rows": (c_int, [_P, _P, c_int64, c_int64, _P, _P]),
    "grpo_logprob_from_logits": (c_int, [_P, c_int, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P]),
    "grpo_logprob_from_logits_bwd": (c_int, [_P, c_int, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, _P, c_int64, _P]),
    "grpo_debug_gemm": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_int, _P]),
    "grpo_debug_plan_units": (c_int, [c_int64, c_int64, c_int, c_int, _P, c_int64, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[ctypes.CDLL] = None


class GrpoLibraryError(RuntimeError):
    """Raised when the CUDA library is missing or an entry point reports failure."""


def load() -> ctypes.CDLL:
    """Open the shared library (once) and declare every prototype. Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GrpoLibraryError(
            f"{LIB_PATH} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU or PyTorch fallback for this path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    """Turn a non-zero ABI return code into an exception (ValueError for argument errors, like the reference)."""
    if rc == 0:
        return
    msg = load().grpo_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(f"{what}: {msg}")
    raise GrpoLibraryError(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def mask_dtype_code(mask) -> int:
    import torch

    if mask is None:
        return MASK_NONE
    if mask.dtype == torch.float32:
        return MASK_F32
    if mask.dtype == torch.int64:
        return MASK_I64
    if mask.dtype in (torch.bool, torch.uint8):
        return MASK_U8
    raise ValueError(f"unsupported mask dtype {mask.dtype}; use float32, int64, bool or uint8")
