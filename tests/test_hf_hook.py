"""The model-side hook must hand the fused head exactly the rows whose logits the reference slices out."""
import pytest
import torch


def test_response_hidden_states_match_logit_slice():
    transformers = pytest.importorskip("transformers")
    from spatialthinker_b200 import hf_hook

    cfg = transformers.Qwen2Config(vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                   num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=128)
    torch.manual_seed(0)
    model = transformers.Qwen2ForCausalLM(cfg).eval()
    bsz, prompt, t_len = 3, 9, 7
    ids = torch.randint(0, 512, (bsz, prompt + t_len))
    mask = torch.ones_like(ids)
    mask[0, :3] = 0  # left padding, as the reference's collate produces
    pos = (mask.cumsum(-1) - 1).clamp(min=0)
    with torch.no_grad():
        logits = model(input_ids=ids, attention_mask=mask, position_ids=pos, use_cache=False).logits
        hidden = hf_hook.response_hidden_states(model, ids, mask, pos, t_len)
        w = hf_hook.lm_head_weight(model)
    assert hidden.shape == (bsz, t_len, 64) and w.shape == (512, 64)
    want = logits[:, -t_len - 1: -1]  # dp_actor.py:150
    got = torch.nn.functional.linear(hidden, w)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    fn = hf_hook.make_hidden_fn(model)
    with torch.no_grad():
        h2 = fn({"input_ids": ids, "attention_mask": mask, "position_ids": pos, "responses": ids[:, -t_len:]})
    assert h2.dtype == torch.bfloat16 and h2.shape == hidden.shape
