"""Model-side glue: get the final hidden states that predict the response tokens out of a HuggingFace
causal LM, instead of its logits.

The reference's actor calls the full model and slices the LOGITS (``logits[:, -T-1:-1]``, dp_actor.py:141-151, or the
packed equivalent :118-139), which materialises ``[tokens, vocab]`` for prompt and response alike. The fused head only
needs ``hidden[:, -T-1:-1]`` - ``T`` rows per sequence - and ``lm_head.weight``.

Only the padded (non ``padding_free``) layout is handled here; the packed layout needs the reference's flash-attn varlen
patch (``verl/models/monkey_patch.py``), which is outside this path.
"""
from __future__ import annotations

from typing import Any, Callable, Dict

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
