"""SURVEY §8 f-1 on the GPU: a (tiny, random-init) HuggingFace model end to end through the forward patch -
``hf_hook.patch_model`` -> fused head -> compared with the unpatched model's ``.logits`` -> cross-entropy, which is what
the reference's ``_forward_micro_batch`` computes (verl/workers/actor/dp_actor.py:141-151); gradients into the body's
parameters; the FSDP root-unit case (fsdp_workers.py:242-280); and the actor loop over a real body."""
import os

import pytest
import torch

from oracle import grpo_oracle as O

pytestmark = pytest.mark.gpu
transformers = pytest.importorskip("transformers")
CLIP = dict(clip_ratio_low=0.2, clip_ratio_high=0.3, clip_ratio_dual=3.0)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def tiny_qwen2(tie=False, vocab=1024, hidden=128):
    cfg = transformers.Qwen2Config(vocab_size=vocab, hidden_size=hidden, intermediate_size=256, num_hidden_layers=2,
                                   num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=128,
                                   tie_word_embeddings=tie, initializer_range=0.1)
    torch.manual_seed(0)
    return transformers.Qwen2ForCausalLM(cfg)


def tiny_qwen2_5_vl(vocab=1024, hidden=128):
    text = dict(vocab_size=vocab, hidden_size=hidden, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                num_key_value_heads=2, max_position_embeddings=256, initializer_range=0.1,
                rope_scaling={"type": "mrope", "mrope_section": [4, 6, 6]})
    vision = dict(depth=2, hidden_size=32, intermediate_size=64, num_heads=2, out_hidden_size=hidden, patch_size=14,
                  spatial_merge_size=2, temporal_patch_size=2, window_size=56, fullatt_block_indexes=[1], in_chans=3)
    cfg = transformers.Qwen2_5_VLConfig(text_config=text, vision_config=vision, image_token_id=1000, video_token_id=1001,
                                        vision_start_token_id=1002)
    torch.manual_seed(0)
    return transformers.Qwen2_5_VLForConditionalGeneration(cfg)


def round_params_to_bf16_(model):
    """The path's precision model (BASELINE.json north_star; actor/config.py:57 ``mp_param_dtype = bf16``): the final hidden
    states and the lm_head weight are bf16 VALUES, everything after them is fp32. fp32 leaves holding bf16-representable
    values keep autograd in fp32 for the comparison."""
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    return model


def reference_response_logits(model, t_len, temperature, **inputs):
    """``logits[:, -T-1:-1] / temperature`` as the reference computes them (dp_actor.py:141-151), with the final hidden
    states rounded to bf16 (what a bf16 body emits; the cast is differentiable, straight-through) and the lm_head product
    in fp32."""
    from spatialthinker_b200 import hf_hook

    hidden = hf_hook.transformer_body(model)(**inputs, use_cache=False).last_hidden_state
    rows = hidden[:, -t_len - 1: -1].to(torch.bfloat16).float()
    return torch.nn.functional.linear(rows, hf_hook.lm_head_weight(model).float()) / temperature


def text_batch(bsz, prompt, t_len, vocab, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab - 30, (bsz, prompt + t_len), generator=g)
    mask = torch.ones_like(ids)
    mask[0, :3] = 0  # left padding, as the reference's collate produces
    pos = (mask.cumsum(-1) - 1).clamp(min=0)
    return ids.to(dev), mask.to(dev), pos.to(dev)


def test_qwen2_5_vl_patched_forward_matches_logits_path(dev):
    """The flagship architecture with an image per sequence: ``out.log_probs`` of the patched forward against the
    cross-entropy of the unpatched model's logits (dp_actor.py:141-151), and the gradients it sends into the body."""
    from spatialthinker_b200 import hf_hook

    bsz, prompt, t_len, vocab = 2, 14, 8, 1024
    model = round_params_to_bf16_(tiny_qwen2_5_vl(vocab).to(dev).train())  # fp32 leaves, bf16-representable values
    ids, mask, pos = text_batch(bsz, prompt, t_len, vocab, dev)
    ids[:, 4:8] = 1000  # one 4 x 4-patch image per sequence = 4 merged image tokens
    pos3 = pos.unsqueeze(0).expand(3, bsz, -1).contiguous()
    g = torch.Generator().manual_seed(1)
    extra = {"pixel_values": torch.randn(2 * 16, 3 * 2 * 14 * 14, generator=g).to(dev),
             "image_grid_thw": torch.tensor([[1, 4, 4], [1, 4, 4]], device=dev)}
    responses = ids[:, -t_len:]
    temperature = 0.8
    coef = torch.randn(bsz, t_len, generator=g).to(dev)

    # reference: the unpatched model's logits, sliced and divided exactly like dp_actor.py:148-151, fp32 cross-entropy
    logits = model(input_ids=ids, attention_mask=mask, position_ids=pos3, use_cache=False, **extra).logits
    z = reference_response_logits(model, t_len, temperature, input_ids=ids, attention_mask=mask, position_ids=pos3, **extra)
    # ... which is the model's own .logits path up to the bf16 rounding of the hidden rows
    assert float((z - logits[:, -t_len - 1: -1] / temperature).abs().max()) < 3e-2
    want_lp = O.log_probs_from_logits(z, responses)
    want_ent = O.entropy_from_logits(z)
    ((want_lp * coef).sum() + 0.1 * want_ent.sum()).backward()
    want_grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    model.zero_grad()

    hf_hook.patch_model(model)
    try:
        out = model(input_ids=ids, attention_mask=mask, position_ids=pos3, **extra, responses=responses,
                    temperature=temperature, want_entropy=True)
        assert out.logits is None and out.log_probs.shape == (bsz, t_len)
        assert float((out.log_probs - want_lp).abs().max()) < 2e-3
        assert float((out.entropy - want_ent).abs().max()) < 2e-3
        ((out.log_probs * coef).sum() + 0.1 * out.entropy.sum()).backward()
        got = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
        assert set(got) == set(want_grads)
        for name in ("lm_head.weight", "model.language_model.norm.weight", "model.language_model.layers.0.self_attn.q_proj.weight",
                     "model.language_model.embed_tokens.weight", "model.visual.blocks.0.attn.qkv.weight"):
            assert name in got, (name, sorted(got)[:8])
            assert rel(got[name], want_grads[name]) < 1e-2, name
        total = sum(float((got[n].float() - want_grads[n].float()).square().sum()) for n in got) ** 0.5
        ref = sum(float(want_grads[n].float().square().sum()) for n in got) ** 0.5
        assert total / ref < 1e-2
        # without `responses` the patched forward is the original one
        again = model(input_ids=ids, attention_mask=mask, position_ids=pos3, use_cache=False, **extra).logits
        assert torch.equal(again, logits)
    finally:
        hf_hook.unpatch_model(model)
    assert "forward" not in model.__dict__


def test_patched_forward_fused_grpo_loss_and_padding_free(dev):
    """``grpo=...``: the patched forward returns the micro-batch loss whose backward reaches the body (three GEMM units
    in the head); ``padding_free_mask``: the packed ``[1, total_nnz, H]`` stream of the padding-free branch."""
    from spatialthinker_b200 import hf_hook

    bsz, prompt, t_len, vocab = 4, 10, 12, 1024
    model = round_params_to_bf16_(tiny_qwen2(vocab=vocab).to(dev).train())
    ids, mask, pos = text_batch(bsz, prompt, t_len, vocab, dev, seed=3)
    responses = ids[:, -t_len:]
    g = torch.Generator().manual_seed(4)
    rmask = (torch.arange(t_len)[None] < torch.randint(3, t_len + 1, (bsz, 1), generator=g)).long().to(dev)
    adv = torch.randn(bsz, 1, generator=g).expand(bsz, t_len).contiguous().to(dev)
    z = reference_response_logits(model, t_len, 1.0, input_ids=ids, attention_mask=mask, position_ids=pos)
    lp_ref = O.log_probs_from_logits(z, responses)
    old = O.perturbed_log_probs(lp_ref.detach().cpu(), seed=5, outlier_frac=0.05).to(dev)
    ref_lp = O.perturbed_log_probs(lp_ref.detach().cpu(), seed=6, outlier_frac=0.05).to(dev)
    loss_ref, met_ref = O.micro_batch_loss(lp_ref, old, adv, rmask, ref_lp, kl_penalty="low_var_kl", kl_coef=0.01,
                                           grad_accum=2.0, **CLIP)
    loss_ref.backward()
    want = {n: p.grad.clone() for n, p in model.named_parameters()}
    model.zero_grad()
    hf_hook.patch_model(model)
    try:
        out = model(input_ids=ids, attention_mask=mask, position_ids=pos, responses=responses,
                    grpo=dict(old_log_probs=old, advantages=adv, response_mask=rmask, ref_log_probs=ref_lp,
                              kl_penalty="low_var_kl", kl_coef=0.01, grad_accum=2.0, **CLIP))
        assert abs(float(out.loss) - float(loss_ref)) <= 1e-2 * abs(float(loss_ref))
        assert abs(float(out.metrics["actor/pg_loss"]) - float(met_ref["actor/pg_loss"])) <= 1e-2 * abs(float(met_ref["actor/pg_loss"]))
        assert float((out.log_probs - lp_ref)[rmask.bool()].abs().max()) < 2e-3
        out.loss.backward()
        for n, p in model.named_parameters():
            assert rel(p.grad, want[n]) < 1e-2, n
        # padding-free: the body's output on the packed token stream; here the packed stream is emulated by gathering the
        # padded body output at the attended positions (the varlen attention itself is the reference's own patch)
        body = hf_hook.transformer_body(model)
        with torch.no_grad():
            hidden = body(input_ids=ids, attention_mask=mask, position_ids=pos, use_cache=False).last_hidden_state
        packed = hidden[mask.bool()].unsqueeze(0)
        rows = hf_hook.packed_response_hidden_states(packed, mask, t_len)
        assert torch.equal(rows, hidden[:, -t_len - 1: -1])
    finally:
        hf_hook.unpatch_model(model)


def test_patched_forward_under_fsdp_root_unit(dev):
    """FSDP (world size 1, the reference's mixed precision: bf16 parameters, fp32 reduce; flat parameters,
    use_orig_params=False): ``lm_head.weight`` - TIED to ``embed_tokens`` as on the 3B checkpoints - is a view into the
    root unit's all-gathered flat parameter while the patched forward runs, and its gradient flows back through autograd
    into that flat parameter. Compared with the same FSDP model through the unpatched logits path."""
    import torch.distributed as dist
    from torch.distributed.fsdp import FullyShardedDataParallel as FSDP
    from torch.distributed.fsdp import MixedPrecision
    from torch.distributed.fsdp.wrap import transformer_auto_wrap_policy
    from transformers.models.qwen2.modeling_qwen2 import Qwen2DecoderLayer
    import functools

    from spatialthinker_b200 import hf_hook

    started = False
    if not dist.is_initialized():
        port = 29600 + os.getpid() % 300
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=dev)
        started = True
    try:
        bsz, prompt, t_len, vocab = 3, 9, 10, 1024
        ids, mask, pos = text_batch(bsz, prompt, t_len, vocab, dev, seed=7)
        responses = ids[:, -t_len:]
        coef = torch.randn(bsz, t_len, generator=torch.Generator().manual_seed(8)).to(dev)
        grads = {}
        for patched in (False, True):
            model = tiny_qwen2(tie=True, vocab=vocab)
            assert model.lm_head.weight is model.model.embed_tokens.weight
            policy = functools.partial(transformer_auto_wrap_policy, transformer_layer_cls={Qwen2DecoderLayer})
            fsdp = FSDP(model, auto_wrap_policy=policy, device_id=dev, use_orig_params=False,
                        mixed_precision=MixedPrecision(param_dtype=torch.bfloat16, reduce_dtype=torch.float32,
                                                       buffer_dtype=torch.float32))
            if patched:
                hf_hook.patch_model(model)  # the wrapped module: FSDP's forward calls model.forward inside the root unit
                out = fsdp(input_ids=ids, attention_mask=mask, position_ids=pos, responses=responses, temperature=0.9)
                lp = out.log_probs
            else:
                logits = fsdp(input_ids=ids, attention_mask=mask, position_ids=pos, use_cache=False).logits
                lp = O.log_probs_from_logits(logits[:, -t_len - 1: -1].float() / 0.9, responses)
            (lp * coef).sum().backward()
            # use_orig_params=False: the gradients live on the flat parameters (one for the root unit - embed_tokens /
            # lm_head (tied) + final norm - and one per decoder layer); both models are built identically, so the flat
            # layouts agree name by name
            grads[patched] = {n: p.grad.detach().float().clone() for n, p in fsdp.named_parameters() if p.grad is not None}
            grads[("lp", patched)] = lp.detach()
        assert float((grads[("lp", True)] - grads[("lp", False)]).abs().max()) < 2e-2  # the unpatched path rounds logits to bf16
        assert set(grads[True]) == set(grads[False]) and len(grads[True]) == 3
        for n in grads[True]:
            assert rel(grads[True][n], grads[False][n]) < 3e-2, n  # both bodies run in bf16; the unpatched head too
    finally:
        if started:
            dist.destroy_process_group()


def test_actor_update_policy_with_hf_body(dev):
    """The actor loop over a real HF body: hidden_fn = hf_hook.make_hidden_fn(model); the head weight is the model's own
    lm_head parameter, the optimizer holds every parameter; one update must move the body and the head and report the
    global gradient norm the unpatched logits path gives."""
    import spatialthinker_b200 as st
    from spatialthinker_b200 import hf_hook

    bsz, prompt, t_len, vocab = 4, 8, 8, 1024
    model = tiny_qwen2(vocab=vocab).to(dev).to(torch.bfloat16).train()
    ids, mask, pos = text_batch(bsz, prompt, t_len, vocab, dev, seed=11)
    responses = ids[:, -t_len:]
    g = torch.Generator().manual_seed(12)
    adv = torch.randn(bsz, 1, generator=g).expand(bsz, t_len).contiguous().to(dev)
    with torch.no_grad():
        logits = model(input_ids=ids, attention_mask=mask, position_ids=pos, use_cache=False).logits
    lp0 = O.log_probs_from_logits(logits[:, -t_len - 1: -1].float(), responses)
    old = O.perturbed_log_probs(lp0.cpu(), seed=13).to(dev)
    # reference gradient norm: logits path, two micro-batches of 2, GA = 2
    for sl in (slice(0, 2), slice(2, 4)):
        lg = model(input_ids=ids[sl], attention_mask=mask[sl], position_ids=pos[sl], use_cache=False).logits
        lp = O.log_probs_from_logits(lg[:, -t_len - 1: -1].float(), responses[sl])
        loss, _ = O.micro_batch_loss(lp, old[sl], adv[sl], mask[sl, -t_len:], None, grad_accum=2.0, **CLIP)
        loss.backward()
    want_norm = float(torch.nn.utils.clip_grad_norm_(model.parameters(), 1e9))
    model.zero_grad()
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    opt = torch.optim.SGD(model.parameters(), lr=0.5)
    cfg = st.ActorConfig(global_batch_size_per_device=4, micro_batch_size_per_device_for_update=2, max_grad_norm=1e9)
    actor = st.DataParallelPPOActor(cfg, hf_hook.lm_head_weight(model), actor_optimizer=opt,
                                    hidden_fn=hf_hook.make_hidden_fn(model))
    data = st.TensorBatch({"input_ids": ids, "attention_mask": mask, "position_ids": pos, "responses": responses,
                           "old_log_probs": old, "advantages": adv}, meta_info={"temperature": 1.0})
    met = actor.update_policy(data)
    assert abs(met["actor/grad_norm"][0] - want_norm) <= 3e-2 * want_norm
    moved = [n for n, p in model.named_parameters() if not torch.equal(p.detach(), before[n])]
    assert "lm_head.weight" in moved and any(n.startswith("model.layers.0") for n in moved)
    assert all(p.grad is None for p in model.parameters())
