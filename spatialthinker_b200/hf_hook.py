"""Model-side glue: get the final hidden states that predict the response tokens out of a HuggingFace
causal LM, instead of its logits.

The reference's actor calls the full model and slices the LOGITS (``logits[:, -T-1:-1]``, dp_actor.py:141-151, or the
packed equivalent :118-139), which materialises ``[tokens, vocab]`` for prompt and response alike. The fused head only
needs ``hidden[:, -T-1:-1]`` - ``T`` rows per sequence - and ``lm_head.weight``.

Two layouts, as in the reference: padded ``[bsz, seqlen, H]`` (:func:`response_hidden_states`) and ``padding_free``
packed ``[total_nnz, H]`` (:func:`packed_response_hidden_states`, dp_actor.py:85-139: the body runs on the unpadded token
stream through the reference's flash-attn varlen patch, which stays the reference's - ``verl/models/monkey_patch.py``).
"""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Dict, Optional

import torch


def transformer_body(model: torch.nn.Module) -> torch.nn.Module:
    """The module that maps tokens to final hidden states (``model.model`` for ``*ForCausalLM`` /
    ``*ForConditionalGeneration``)."""
    body = getattr(model, "model", None)
    if body is None or not isinstance(body, torch.nn.Module):
        raise TypeError(f"{type(model).__name__} has no `.model` transformer body")
    return body


def lm_head_weight(model: torch.nn.Module) -> torch.Tensor:
    """``[vocab, hidden]`` output-projection weight (tied to the embedding on the 3B checkpoints)."""
    head = model.get_output_embeddings() if hasattr(model, "get_output_embeddings") else getattr(model, "lm_head", None)
    if head is None or getattr(head, "bias", None) is not None:
        raise TypeError("expected a bias-free lm_head (nn.Linear(hidden, vocab, bias=False))")
    return head.weight


def response_hidden_states(model: torch.nn.Module, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                           position_ids: torch.Tensor, response_length: int, **model_kwargs: Any) -> torch.Tensor:
    """Run the transformer body and return ``hidden[:, -T-1:-1]``: row ``t`` predicts ``responses[:, t]``
    (the slice the reference applies to the logits, dp_actor.py:150)."""
    if position_ids is not None and position_ids.dim() == 3:  # qwen2-vl mrope: (bsz, 3, seqlen) -> (3, bsz, seqlen)
        position_ids = position_ids.transpose(0, 1)
    out = transformer_body(model)(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                                  use_cache=False, **model_kwargs)
    hidden = out.last_hidden_state if hasattr(out, "last_hidden_state") else out[0]
    return hidden[:, -response_length - 1: -1]


class _PackedResponseRows(torch.autograd.Function):
    """Rows of a packed ``[total_nnz, H]`` hidden-state stream that predict the response tokens, as ``[bsz, T, H]``.

    The reference pads the per-token LOG-PROBS back to ``(bsz, seqlen)`` (``pad_input``, dp_actor.py:134-138) and slices
    ``[:, -T-1:-1]`` (:139); here the same index map is applied to the HIDDEN rows before the head, so only ``bsz * T``
    rows reach it. Index map and row moves are this library's kernels (``grpo_compact_index`` / ``grpo_scatter_rows``);
    positions whose ``attention_mask`` is 0 give zero rows (their ``response_mask`` is 0 too), nothing synchronises.
    """

    @staticmethod
    def forward(ctx, hidden_packed, attention_mask, response_length: int):
        from .fused import compact_index, scatter_rows

        bsz, seqlen = attention_mask.shape
        t_len = int(response_length)
        if not 0 < t_len < seqlen:
            raise ValueError(f"response_length {t_len} does not fit a sequence length of {seqlen}")
        gather_idx, inverse, _ = compact_index(attention_mask)  # packed <-> (bsz * seqlen) slot maps
        rows = inverse.view(bsz, seqlen)[:, seqlen - t_len - 1: seqlen - 1].contiguous().view(-1)
        out = scatter_rows(hidden_packed, rows)  # out[slot] = hidden_packed[rows[slot]] or 0
        ctx.save_for_backward(gather_idx)
        ctx.dims = (bsz, seqlen, t_len, hidden_packed.shape[0])
        return out.view(bsz, t_len, hidden_packed.shape[-1])

    @staticmethod
    def backward(ctx, grad):
        from .fused import scatter_rows

        (gather_idx,) = ctx.saved_tensors
        bsz, seqlen, t_len, nnz = ctx.dims
        pos = gather_idx[:nnz].long()  # flattened (sequence, position) of every packed token
        seq = pos // seqlen
        t = pos - seq * seqlen - (seqlen - t_len - 1)
        slot = torch.where((t >= 0) & (t < t_len), seq * t_len + t, torch.full_like(t, -1)).to(torch.int32)
        d_packed = scatter_rows(grad.reshape(bsz * t_len, -1), slot)  # every packed row is read by at most one slot
        return d_packed, None, None


def packed_response_hidden_states(hidden_packed: torch.Tensor, attention_mask: torch.Tensor,
                                  response_length: int) -> torch.Tensor:
    """``padding_free`` layout: ``hidden_packed`` is the body's output on the unpadded token stream
    (``[total_nnz, H]`` or ``[1, total_nnz, H]``, tokens in ``attention_mask`` order as ``unpad_input`` produces them,
    dp_actor.py:86-89). Returns ``[bsz, T, H]`` with row ``t`` predicting ``responses[:, t]``; differentiable."""
    if hidden_packed.dim() == 3:
        hidden_packed = hidden_packed.squeeze(0)
    return _PackedResponseRows.apply(hidden_packed.contiguous(), attention_mask, response_length)


def make_hidden_fn(model: torch.nn.Module) -> Callable[[Dict[str, Any]], torch.Tensor]:
    """``hidden_fn`` for :class:`spatialthinker_b200.dp_actor.DataParallelPPOActor`: micro-batch dict -> hidden states."""

    def hidden_fn(micro: Dict[str, Any]) -> torch.Tensor:
        extra = {}
        if "multi_modal_inputs" in micro:  # same concatenation as dp_actor.py:78-83
            for key in micro["multi_modal_inputs"][0].keys():
                extra[key] = torch.cat([inputs[key] for inputs in micro["multi_modal_inputs"]], dim=0)
        return response_hidden_states(model, micro["input_ids"], micro["attention_mask"], micro["position_ids"],
                                      micro["responses"].size(-1), **extra).to(torch.bfloat16)

    return hidden_fn


# ----------------------------------------------------------------------------------------------------------------
# forward patch: the model returns .log_probs instead of .logits
# ----------------------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class FusedHeadOutput:
    """What a patched ``forward`` returns when it is given ``responses``: per-token log-probs of the response tokens
    (and optionally their entropy, or the fused GRPO loss) instead of ``[bsz, seqlen, vocab]`` logits.

    A dataclass on purpose: FSDP and DDP walk a module's outputs (``torch.distributed.utils._apply_to_tensors``: tensors,
    dicts, tuples, dataclasses) to hang their pre-backward hooks on every tensor; an opaque object would leave the root
    unit resharded when the head's backward runs."""

    log_probs: torch.Tensor
    entropy: Optional[torch.Tensor] = None
    loss: Optional[torch.Tensor] = None
    metrics: Optional[Dict[str, Any]] = None
    last_hidden_state: Optional[torch.Tensor] = None
    logits: None = None  # never materialised


_PATCH_ATTR = "_grpo_b200_original_forward"


def patch_model(model: torch.nn.Module) -> torch.nn.Module:
    """Swap ``model.forward`` (``Qwen2_5_VLForConditionalGeneration``, ``Qwen2VLForConditionalGeneration``,
    ``Qwen2ForCausalLM`` and any other ``*ForCausalLM`` whose body is ``model.model`` and whose head is a bias-free
    ``lm_head``) for one that can return log-probs. The sibling of the reference's attention patch
    (verl/models/monkey_patch.py:22-32); the call site it serves is ``_forward_micro_batch`` (dp_actor.py:64-153).

    The patched forward behaves exactly like the original unless it is called with ``responses=`` (``[bsz, T]`` int64):
    then it runs the transformer body only, keeps the ``T`` hidden rows per sequence that predict the response tokens
    (``hidden[:, -T-1:-1]``, dp_actor.py:150; or, with ``padding_free_mask=attention_mask``, the same rows out of the
    packed ``[1, total_nnz, H]`` stream of the padding-free branch, dp_actor.py:118-139) and feeds them to the fused head
    together with ``self.lm_head.weight``. Extra keyword arguments:

    * ``temperature`` (float, default 1.0), ``want_entropy`` (bool) -> ``out.log_probs`` / ``out.entropy`` ``[bsz, T]``,
      differentiable (autograd into the body and into ``lm_head.weight``; the backward recomputes the logits tiles);
    * ``grpo=dict(old_log_probs=, advantages=, response_mask=, ref_log_probs=None, clip_ratio_low=, ..., kl_penalty=,
      kl_coef=, grad_accum=)`` -> additionally ``out.loss`` (scalar to call ``.backward()`` on; the gradients are produced
      in the forward pass, three GEMM units in total) and ``out.metrics`` (the ``actor/*`` scalars, on the device).

    FSDP: the head runs INSIDE the wrapped module's forward, i.e. while the root FSDP unit - which owns ``embed_tokens``,
    the final norm and ``lm_head`` under the reference's wrap policy (fsdp_workers.py:242-280) - has its flat parameter
    all-gathered, so ``self.lm_head.weight`` is the full ``[V, H]`` view; its gradient flows back through autograd into
    the flat parameter and is reduce-scattered by FSDP like any other (a weight tied to ``embed_tokens``, as on the 3B
    checkpoints, simply receives both contributions). Nothing needs ``summon_full_params``.
    """
    if getattr(model, _PATCH_ATTR, None) is not None:
        return model
    transformer_body(model)  # raises for unsupported architectures
    lm_head_weight(model)
    original = model.forward

    def forward(*args, responses=None, temperature: float = 1.0, want_entropy: bool = False, grpo=None,
                padding_free_mask=None, **kwargs):
        if responses is None:
            return original(*args, **kwargs)
        from .fused import fused_grpo_loss, fused_lm_head_log_probs

        kwargs.pop("labels", None)
        kwargs.setdefault("use_cache", False)
        out = transformer_body(model)(*args, **kwargs)
        hidden = out.last_hidden_state if hasattr(out, "last_hidden_state") else out[0]
        t_len = responses.size(-1)
        if padding_free_mask is not None:
            rows = packed_response_hidden_states(hidden, padding_free_mask, t_len)
        else:
            rows = hidden[:, -t_len - 1: -1]
        rows = rows.to(torch.bfloat16)
        weight = lm_head_weight(model)
        if weight.dtype != torch.bfloat16:
            weight = weight.to(torch.bfloat16)  # differentiable cast (fp32 master weights)
        if grpo is None:
            logp, ent = fused_lm_head_log_probs(rows, weight, responses, temperature, want_entropy)
            return FusedHeadOutput(log_probs=logp, entropy=ent, last_hidden_state=hidden)
        g = dict(grpo)
        loss, metrics = fused_grpo_loss(rows, weight, responses, g.pop("old_log_probs"), g.pop("advantages"),
                                        g.pop("ref_log_probs", None), g.pop("response_mask"), temperature=temperature,
                                        want_entropy=want_entropy, **g)
        return FusedHeadOutput(log_probs=metrics.pop("log_probs"), entropy=metrics.pop("entropy", None), loss=loss,
                               metrics=metrics, last_hidden_state=hidden)

    setattr(model, _PATCH_ATTR, original)
    model.forward = forward
    return model


def unpatch_model(model: torch.nn.Module) -> torch.nn.Module:
    """Undo :func:`patch_model`."""
    original = getattr(model, _PATCH_ATTR, None)
    if original is None:
        return model
    model.__dict__.pop("forward", None)  # the patched function lives in the instance dict
    if getattr(original, "__func__", None) is not type(model).forward:  # forward had already been overridden per instance
        model.forward = original
    delattr(model, _PATCH_ATTR)
    return model
