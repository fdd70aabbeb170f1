"""Monkey-patch a live veRL / EasyR1 checkout so its call sites run on this library.

The reference binds its ops through module attributes - ``verl.utils.torch_functional`` (imported as ``VF`` at
dp_actor.py:29 and ray_trainer.py) and ``verl.trainer.core_algos`` (dp_actor.py:28, ray_trainer.py:158) - so replacing
the attributes is enough for everything except ``DataParallelPPOActor.log_probs_from_logits``, which is bound at
construction (dp_actor.py:59-62): call :func:`patch_verl` before the actor is built.
"""
from __future__ import annotations

import importlib
from typing import Dict, Tuple

_SAVED: Dict[Tuple[str, str], object] = {}


def _device_dispatch(orig, mine):
    """CUDA tensors run on this library. The reference also calls these functions on the CPU driver
    (ray_trainer.py:132,158): there, tensors are staged through the visible GPU; a driver process without any GPU keeps
    the reference's own function (that is the reference running, not a fallback inside this package)."""
    import torch

    def wrapper(*args, **kwargs):
        tensors = [a for a in list(args) + list(kwargs.values()) if isinstance(a, torch.Tensor)]
        if tensors and all(t.is_cuda for t in tensors):
            return mine(*args, **kwargs)
        if not torch.cuda.is_available():
            return orig(*args, **kwargs)
        up = lambda a: a.cuda(non_blocking=True) if isinstance(a, torch.Tensor) else a  # noqa: E731
        out = mine(*[up(a) for a in args], **{k: up(v) for k, v in kwargs.items()})
        if isinstance(out, tuple):
            first = out[0].cpu()
            return tuple(first if o is out[0] else o.cpu() for o in out)  # keeps `returns is advantages`
        return out.cpu()

    wrapper.__wrapped__ = orig
    wrapper.__name__ = getattr(orig, "__name__", "patched")
    return wrapper


def patch_verl(*, advantages: bool = True, loss: bool = True, log_probs: bool = True) -> Dict[str, list]:
    """Replace the reference's hot-path functions in place. Returns the names that were patched per module."""
    from . import core_algos as my_ca
    from . import torch_functional as my_vf

    done: Dict[str, list] = {}
    vf = importlib.import_module("verl.utils.torch_functional")
    ca = importlib.import_module("verl.trainer.core_algos")
    plan = []
    if log_probs:
        plan += [(vf, "log_probs_from_logits", my_vf.log_probs_from_logits), (vf, "masked_mean", my_vf.masked_mean)]
    if advantages:
        plan += [(ca, name, getattr(my_ca, name)) for name in (
            "compute_grpo_outcome_advantage", "compute_rloo_outcome_advantage", "compute_remax_outcome_advantage",
            "compute_reinforce_plus_plus_outcome_advantage", "compute_gae_advantage_return")]
        plan += [(vf, "masked_whiten", my_vf.masked_whiten)]
    if loss:
        plan += [(ca, "compute_policy_loss", my_ca.compute_policy_loss), (ca, "compute_kl", my_ca.compute_kl),
                 (ca, "compute_value_loss", my_ca.compute_value_loss)]
    for mod, name, fn in plan:
        key = (mod.__name__, name)
        if key not in _SAVED:
            _SAVED[key] = getattr(mod, name)
        setattr(mod, name, _device_dispatch(_SAVED[key], fn))
        done.setdefault(mod.__name__, []).append(name)
    return done


def unpatch_verl() -> None:
    for (mod_name, name), fn in list(_SAVED.items()):
        setattr(importlib.import_module(mod_name), name, fn)
        del _SAVED[(mod_name, name)]
