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


def test_qwen2_5_vl_multimodal_hidden_fn():
    """The flagship architecture (Qwen2.5-VL, tiny random-init config): mrope position ids arrive as (bsz, 3, seqlen) and
    the images as a per-sample list of processor outputs, exactly what the reference's actor handles in
    dp_actor.py:72-83; the hook must hand back the rows whose logits the reference slices (:150)."""
    transformers = pytest.importorskip("transformers")
    if not hasattr(transformers, "Qwen2_5_VLForConditionalGeneration"):
        pytest.skip("transformers without Qwen2.5-VL")
    from spatialthinker_b200 import hf_hook

    text = dict(vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                num_key_value_heads=2, max_position_embeddings=256,
                rope_scaling={"type": "mrope", "mrope_section": [2, 3, 3]})
    vision = dict(depth=2, hidden_size=32, intermediate_size=64, num_heads=2, out_hidden_size=64, patch_size=14,
                  spatial_merge_size=2, temporal_patch_size=2, window_size=56, fullatt_block_indexes=[1], in_chans=3)
    cfg = transformers.Qwen2_5_VLConfig(text_config=text, vision_config=vision, image_token_id=500, video_token_id=501,
                                        vision_start_token_id=502)
    torch.manual_seed(0)
    model = transformers.Qwen2_5_VLForConditionalGeneration(cfg).eval()
    bsz, prompt, t_len = 2, 12, 6
    ids = torch.randint(0, 400, (bsz, prompt + t_len))
    ids[:, 4:8] = 500  # one 4 x 4-patch image per sequence = 4 merged image tokens
    mask = torch.ones_like(ids)
    mask[0, :3] = 0
    mm = [{"pixel_values": torch.randn(16, 3 * 2 * 14 * 14), "image_grid_thw": torch.tensor([[1, 4, 4]])}
          for _ in range(bsz)]
    pos = (mask.cumsum(-1) - 1).clamp(min=0)
    pos_b3 = pos.unsqueeze(1).expand(bsz, 3, -1).contiguous()  # the reference's batch layout (bsz, 3, seqlen)
    extra = {k: torch.cat([m[k] for m in mm], dim=0) for k in mm[0]}
    with torch.no_grad():
        logits = model(input_ids=ids, attention_mask=mask, position_ids=pos_b3.transpose(0, 1), use_cache=False,
                       **extra).logits
        hidden = hf_hook.response_hidden_states(model, ids, mask, pos_b3, t_len, **extra)
        w = hf_hook.lm_head_weight(model)
        torch.testing.assert_close(torch.nn.functional.linear(hidden, w), logits[:, -t_len - 1: -1], rtol=1e-5, atol=1e-5)
        h2 = hf_hook.make_hidden_fn(model)({"input_ids": ids, "attention_mask": mask, "position_ids": pos_b3,
                                            "responses": ids[:, -t_len:], "multi_modal_inputs": mm})
    assert h2.dtype == torch.bfloat16 and h2.shape == (bsz, t_len, 64)
    torch.testing.assert_close(h2.float(), hidden, rtol=1e-2, atol=1e-2)
    # the image really reaches the body: other pixels, other hidden states
    mm2 = [{**m, "pixel_values": m["pixel_values"] + 1.0} for m in mm]
    with torch.no_grad():
        h3 = hf_hook.make_hidden_fn(model)({"input_ids": ids, "attention_mask": mask, "position_ids": pos_b3,
                                            "responses": ids[:, -t_len:], "multi_modal_inputs": mm2})
    assert not torch.equal(h3, h2)
