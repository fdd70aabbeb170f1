"""spatialthinker_b200 - a B200-native (sm_100a) GRPO policy-loss path for SpatialThinker's veRL actor.

Host side: Python/PyTorch mirrors of the reference's call surface
(:mod:`torch_functional`, :mod:`core_algos`, :mod:`dp_actor`) plus the fused entry points (:mod:`fused`).
Device side: hand-written CUDA behind a C ABI (``include/grpo_b200.h`` -> ``libgrpo_b200.so``).
There is no CPU fallback: every function raises if its CUDA library or a CUDA tensor is missing.
"""
from . import core_algos, dp_actor, fused, hf_hook, ray_trainer, sharding, torch_functional  # noqa: F401
from ._lib import GrpoLibraryError, load as load_library  # noqa: F401
from .core_algos import (  # noqa: F401
    compute_gae_advantage_return,
    compute_grpo_outcome_advantage,
    compute_kl,
    compute_policy_loss,
    compute_reinforce_plus_plus_outcome_advantage,
    compute_remax_outcome_advantage,
    compute_rewards,
    compute_rloo_outcome_advantage,
    compute_value_loss,
    kl_penalty,
)
from .dp_actor import ActorConfig, DataParallelPPOActor  # noqa: F401
from .fused import DeferredDW, fused_grpo_loss, fused_lm_head_log_probs, grpo_micro_batch_step  # noqa: F401
from .patch import patch_verl, unpatch_verl  # noqa: F401
from .protocol import TensorBatch  # noqa: F401
from .ray_trainer import AdvantageEstimator, apply_kl_penalty, compute_advantage, experience_pass  # noqa: F401
from .torch_functional import (  # noqa: F401
    entropy_from_logits,
    log_probs_from_logits,
    logprobs_from_logits,
    masked_mean,
    masked_var,
    masked_whiten,
)

__version__ = "0.1.0"
