"""GPU, model level: a random-init Llama patched with KVCompress runs generate(); its
decode logits are compared with the same model running on the oracle's dense cache (reference semantics)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model():
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=4, num_attention_heads=8,
                      num_key_value_heads=2, head_dim=64, vocab_size=512, max_position_embeddings=4096,
                      rope_theta=500000.0)
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(0)
    return LlamaForCausalLM(cfg).to(device="cuda", dtype=torch.bfloat16).eval()


@torch.no_grad()
def _decode_logits(model, cache, ids, steps):
    out = model(input_ids=ids, past_key_values=cache, use_cache=True)
    logits = [out.logits[:, -1].float()]
    tok = out.logits[:, -1].argmax(-1, keepdim=True)
    toks = [tok]
    for _ in range(steps):
        out = model(input_ids=tok, past_key_values=cache, use_cache=True)
        logits.append(out.logits[:, -1].float())
        tok = out.logits[:, -1].argmax(-1, keepdim=True)
        toks.append(tok)
    return torch.stack(logits), torch.cat(toks, 1)


def test_patched_llama_decode_matches_oracle_cache():
    from tests.oracle_cache import OracleCache
    from xkv_b200.configurations import generate_consecutive_xKV_config
    from xkv_b200.customized_cache import FakeLayerMergingCache
    from xkv_b200.patch import KVCompress

    model = _model()
    cfg = generate_consecutive_xKV_config(num_layers=4, end_layer=-1, group_size=2, rank_k=64, rank_v=128)
    KVCompress(xKV_config=cfg)(model)
    ids = torch.randint(0, 512, (1, 700), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    lg_ours, tk_ours = _decode_logits(model, FakeLayerMergingCache(cfg), ids, steps=6)
    for layer in model.model.layers:
        layer.self_attn.xkv_fused_decode = False
    lg_ref, tk_ref = _decode_logits(model, OracleCache(cfg), ids, steps=6)
    torch.cuda.synchronize()
    # prefill logits do not depend on the cache at all (prefill attends on the original KV, llama.py:46-50)
    assert torch.allclose(lg_ours[0], lg_ref[0], atol=1e-3)
    dev = (lg_ours[1:] - lg_ref[1:]).abs().max().item()
    scale = lg_ref[1:].abs().max().item()
    print(f"decode logits: max |ours - oracle| = {dev:.4f} (logit scale {scale:.3f})")
    assert dev <= 5e-2 * scale


def test_generate_through_the_patch():
    from xkv_b200 import ops
    from xkv_b200.configurations import generate_consecutive_xKV_config
    from xkv_b200.patch import KVCompress

    model = _model()
    cfg = generate_consecutive_xKV_config(num_layers=4, end_layer=-1, group_size=4, rank_k=64, rank_v=64)
    KVCompress(xKV_config=cfg)(model)
    ids = torch.randint(0, 512, (1, 300), device="cuda")
    before = ops.launch_count()
    out = model.generate(ids, max_new_tokens=5, do_sample=False)
    torch.cuda.synchronize()
    assert out.shape == (1, 305)
    assert ops.launch_count() > before      # the CUDA path ran (no silent eager fallback)


def _mla_model():
    from transformers import DeepseekV2Config, DeepseekV2ForCausalLM

    cfg = DeepseekV2Config(hidden_size=256, intermediate_size=512, moe_intermediate_size=128, num_hidden_layers=3,
                           num_attention_heads=4, num_key_value_heads=4, kv_lora_rank=512, q_lora_rank=None,
                           qk_nope_head_dim=64, qk_rope_head_dim=64, v_head_dim=64, head_dim=64, vocab_size=512,
                           n_routed_experts=4, n_shared_experts=1, num_experts_per_tok=2, first_k_dense_replace=3,
                           max_position_embeddings=4096)
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(0)
    return DeepseekV2ForCausalLM(cfg).to(device="cuda", dtype=torch.bfloat16).eval()


def test_patched_deepseek_v2_mla_matches_oracle_cache():
    """latent-slot caching (kv_lora_rank 512 in the key slot, k_pe in the value slot,
    re_apply_rope=False, merge_value=False), prefill AND decode see the cache's return."""
    from tests.oracle_cache import OracleCache
    from xkv_b200.configurations import generate_consecutive_xKV_config
    from xkv_b200.customized_cache import FakeLayerMergingCache
    from xkv_b200.patch import KVCompress

    model = _mla_model()
    cfg = generate_consecutive_xKV_config(num_layers=3, end_layer=-1, group_size=3, rank_k=256, rank_v=None,
                                          merge_value=False)
    KVCompress(xKV_config=cfg)(model)
    ids = torch.randint(0, 512, (1, 600), device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    lg_ours, _ = _decode_logits(model, FakeLayerMergingCache(cfg), ids, steps=4)
    lg_ref, _ = _decode_logits(model, OracleCache(cfg), ids, steps=4)
    torch.cuda.synchronize()
    dev = (lg_ours - lg_ref).abs().max().item()
    scale = lg_ref.abs().max().item()
    print(f"MLA logits: max |ours - oracle| = {dev:.4f} (logit scale {scale:.3f})")
    assert dev <= 5e-2 * scale
    bad = generate_consecutive_xKV_config(num_layers=3, end_layer=-1, group_size=3, rank_k=256, rank_v=64)
    with pytest.raises(ValueError, match="merge_v"):
        model(input_ids=ids[:, :300], past_key_values=FakeLayerMergingCache(bad), use_cache=True)
