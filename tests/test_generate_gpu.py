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


def test_llama31_8b_geometry_xkv4_long_prompt():
    """Config 2's geometry through the real patch path: hidden 4096, 32 query / 8 kv heads x 128 (GQA 4), RoPE theta
    500000, xKV-4 with rank 512 / 768, an 8192-token prompt; 8 layers (two groups) and a narrow MLP keep the random-init
    model small.  Decode runs the default fused kernel (right-factor slice in shared memory, score MMA); the oracle
    cache runs the reference's arithmetic (torch SVD of the 8192 x 4096 group matrices) on the same device.

    The prompt draws from 48 distinct tokens so that the KV of a RANDOM-INIT model is compressible at all (rank-512
    error ~3 %): with i.i.d. tokens its spectrum is flat, both caches lose 62 % of K and the two (equally good, 0.2 %
    apart in error) subspaces give unrelated logits -- measured, tools/diag_generate.py.  Parity is asserted (i) on the
    reconstruction error of the stored cache against the reference's at equal rank (the north-star criterion, here
    through the whole model path) and (ii) on the decode logits."""
    from transformers import LlamaConfig, LlamaForCausalLM

    from tests.oracle_cache import OracleCache
    from xkv_b200.configurations import generate_consecutive_xKV_config
    from xkv_b200.customized_cache import FakeLayerMergingCache
    from xkv_b200.patch import KVCompress

    mc = LlamaConfig(hidden_size=4096, intermediate_size=1024, num_hidden_layers=8, num_attention_heads=32,
                     num_key_value_heads=8, head_dim=128, vocab_size=1024, max_position_embeddings=16384,
                     rope_theta=500000.0)
    mc._attn_implementation = "sdpa"
    torch.manual_seed(0)
    model = LlamaForCausalLM(mc).to(device="cuda", dtype=torch.bfloat16).eval()
    cfg = generate_consecutive_xKV_config(num_layers=8, end_layer=-1, group_size=4, rank_k=512, rank_v=768)
    KVCompress(xKV_config=cfg)(model)
    S = 8192
    ids = torch.randint(0, 48, (1, S), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    cache = FakeLayerMergingCache(cfg)
    lg_ours, _ = _decode_logits(model, cache, ids, steps=4)
    assert all(cache.layers[i].group is not None for i in range(8))          # every layer ended up factored
    assert cache.layers[0].group.factors.key.A.shape == (S, 512)
    assert cache.layers[0].group.factors.value.A.shape == (S, 768)
    for layer in model.model.layers:
        layer.self_attn.xkv_fused_decode = False
    oracle = OracleCache(cfg)
    lg_ref, _ = _decode_logits(model, oracle, ids, steps=4)
    dense = OracleCache(generate_consecutive_xKV_config(num_layers=8, end_layer=-1, group_size=4, rank_k=10 ** 6,
                                                        rank_v=10 ** 6))     # rank >= min(m, n): the reference's no-op
    lg_dense, _ = _decode_logits(model, dense, ids, steps=4)
    torch.cuda.synchronize()
    assert torch.allclose(lg_ours[0], lg_ref[0], atol=1e-3)

    def rel(a, b):
        return ((a.float() - b.float()).norm() / b.float().norm()).item()

    for li in (0, 5):
        k_o, v_o = cache.materialize(li)
        for name, ours, ref, exact in (("K", k_o, oracle.layers[li].keys, dense.layers[li].keys),
                                       ("V", v_o, oracle.layers[li].values, dense.layers[li].values)):
            e_ours, e_ref = rel(ours[:, :, :S], exact[:, :, :S]), rel(ref[:, :, :S], exact[:, :, :S])
            print(f"layer {li} {name}: stored-cache error ours {e_ours:.5f} reference {e_ref:.5f}")
            # per-layer slices of a group's error (the group total is what is within 1 %): 2 % + the bf16 storage floor
            assert e_ours ** 2 <= (1.02 * e_ref) ** 2 + 3e-3 ** 2
    dev = (lg_ours[1:] - lg_ref[1:]).abs().max().item()
    scale = lg_ref[1:].abs().max().item()
    loss = (lg_ref[1:] - lg_dense[1:]).abs().max().item()
    print(f"Llama-3.1-8B geometry, {S}-token prompt: decode logits max |ours - oracle| = {dev:.4f} "
          f"(logit scale {scale:.3f}; oracle vs uncompressed {loss:.4f})")
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
    from xkv_b200 import ops

    calls = []
    orig = ops.decode_absorbed
    ops.decode_absorbed = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        lg_ours, _ = _decode_logits(model, FakeLayerMergingCache(cfg), ids, steps=6)
    finally:
        ops.decode_absorbed = orig
    assert len(calls) == 3 * 6        # every layer of every decode step ran in the factors' rank space (no latent rebuild)
    for layer in model.model.layers:
        layer.self_attn.xkv_fused_decode = False
    lg_dense, _ = _decode_logits(model, FakeLayerMergingCache(cfg), ids, steps=6)   # materialise + kv_b_proj over the cache
    lg_ref, _ = _decode_logits(model, OracleCache(cfg), ids, steps=6)
    for layer in model.model.layers:
        layer.self_attn.xkv_fused_decode = True
    torch.cuda.synchronize()
    dev = (lg_ours - lg_ref).abs().max().item()
    dev_dense = (lg_dense - lg_ref).abs().max().item()
    scale = lg_ref.abs().max().item()
    print(f"MLA logits: max |absorbed - oracle| = {dev:.4f}, max |materialised - oracle| = {dev_dense:.4f} (logit scale {scale:.3f})")
    assert dev <= 5e-2 * scale and dev_dense <= 5e-2 * scale
    bad = generate_consecutive_xKV_config(num_layers=3, end_layer=-1, group_size=3, rank_k=256, rank_v=64)
    with pytest.raises(ValueError, match="merge_v"):
        model(input_ids=ids[:, :300], past_key_values=FakeLayerMergingCache(bad), use_cache=True)


def _windowed_model(family: str, window: int):
    if family == "mistral":
        from transformers import MistralConfig, MistralForCausalLM

        cfg = MistralConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=4, num_attention_heads=8,
                            num_key_value_heads=2, head_dim=64, vocab_size=512, max_position_embeddings=4096,
                            sliding_window=window)
        cls = MistralForCausalLM
    else:
        from transformers import Qwen2Config, Qwen2ForCausalLM

        # head_dim = 512 / 8 = 64 (the fused kernel handles 64 and 128)
        cfg = Qwen2Config(hidden_size=512, intermediate_size=512, num_hidden_layers=4, num_attention_heads=8,
                          num_key_value_heads=2, vocab_size=512, max_position_embeddings=4096, sliding_window=window,
                          use_sliding_window=True, max_window_layers=2)   # layers 2, 3 slide, layers 0, 1 attend fully
        cls = Qwen2ForCausalLM
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(0)
    return cls(cfg).to(device="cuda", dtype=torch.bfloat16).eval()


@pytest.mark.parametrize("family,window", [("mistral", 256), ("mistral", 4096), ("qwen", 256)])
def test_sliding_window_checkpoints_match_oracle_cache(family, window):
    """Reference mistral.py:69 forwards ``sliding_window`` to SDPA, whose mask hides tokens beyond the window.  The
    fused decode kernel has no window, so a layer whose context outgrew its window must take the dense path (and the
    fused path otherwise): decode logits against the oracle cache with a window SHORTER than the 700-token prompt,
    with one that is longer (fused kernel runs), and on a Qwen2 model where only some layers slide."""
    from tests.oracle_cache import OracleCache
    from xkv_b200 import ops
    from xkv_b200.configurations import generate_consecutive_xKV_config
    from xkv_b200.customized_cache import FakeLayerMergingCache
    from xkv_b200.patch import KVCompress

    model = _windowed_model(family, window)
    cfg = generate_consecutive_xKV_config(num_layers=4, end_layer=-1, group_size=2, rank_k=64, rank_v=128)
    KVCompress(xKV_config=cfg)(model)
    ids = torch.randint(0, 512, (1, 700), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    lg_ours, _ = _decode_logits(model, FakeLayerMergingCache(cfg), ids, steps=4)
    for layer in model.model.layers:
        layer.self_attn.xkv_fused_decode = False
    lg_ref, _ = _decode_logits(model, OracleCache(cfg), ids, steps=4)
    # the same model with the window ignored (what the fused kernel would compute): must differ when the window bites
    torch.cuda.synchronize()
    assert torch.allclose(lg_ours[0], lg_ref[0], atol=1e-3)
    dev = (lg_ours[1:] - lg_ref[1:]).abs().max().item()
    scale = lg_ref[1:].abs().max().item()
    print(f"{family} window {window}: decode logits max |ours - oracle| = {dev:.4f} (logit scale {scale:.3f})")
    assert dev <= 5e-2 * scale
    # which path ran: count fused-decode launches of one more step
    for layer in model.model.layers:
        layer.self_attn.xkv_fused_decode = True
    cache = FakeLayerMergingCache(cfg)
    out = model(input_ids=ids, past_key_values=cache, use_cache=True)
    calls = []
    orig = ops.decode_attention
    ops.decode_attention = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        model(input_ids=out.logits[:, -1].argmax(-1, keepdim=True), past_key_values=cache, use_cache=True)
    finally:
        ops.decode_attention = orig
    expect = {("mistral", 256): 0, ("mistral", 4096): 4, ("qwen", 256): 2}[(family, window)]
    assert len(calls) == expect, f"{len(calls)} layers used the fused kernel, expected {expect}"


@pytest.mark.slow
def test_llama31_8b_full_scale_generate_64k():
    """SURVEY section 8 f1 at the stated scale: a random-init Llama-3.1-8B (32 layers, hidden 4096, 32 / 8 heads x 128,
    intermediate 14336, vocabulary 128256), a 65536-token prompt through ``model.generate`` with the xKV-4 patch
    (rank 512 / 768), 8 decode steps.  Checked: every layer ends up factored at the expected shapes; the stored cache of
    sampled groups is within 1 % of the Eckart-Young optimum (the matrix the reference multiplies back), measured through
    the fp64 Gram of the exact pre-RoPE keys / values captured at prefill; the fused decode logits are finite and agree
    with the same cache read through the dense-compat path (materialise + SDPA)."""
    import gc

    from transformers import LlamaConfig, LlamaForCausalLM

    from xkv_b200.configurations import generate_consecutive_xKV_config
    from xkv_b200.customized_cache import FakeLayerMergingCache
    from xkv_b200.customized_cache import fake_layer_merge_dynamic_cache as cache_mod
    from xkv_b200.patch import KVCompress

    S, L = 65536, 32
    mc = LlamaConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=L, num_attention_heads=32,
                     num_key_value_heads=8, head_dim=128, vocab_size=128256, max_position_embeddings=131072,
                     rope_theta=500000.0)
    mc._attn_implementation = "sdpa"
    torch.manual_seed(0)
    with torch.device("cuda"):
        model = LlamaForCausalLM(mc).to(torch.bfloat16).eval()
    cfg = generate_consecutive_xKV_config(num_layers=L, end_layer=-1, group_size=4, rank_k=512, rank_v=768)
    KVCompress(xKV_config=cfg)(model)
    # 48 distinct tokens: the KV of a random-init model is only compressible when the prompt is (see the 8-layer test)
    ids = torch.randint(0, 48, (1, S), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))

    # capture the exact group matrices of two groups at merge time (what fake_svd would see), as fp64 Grams
    grams = {}
    orig_merge = FakeLayerMergingCache.grouped_layer_merging

    def spy(self, last_layer_idx):
        if last_layer_idx in (3, 31):
            info = self.merge_setup.get_group_for_layer(last_layer_idx)
            for name, attr in (("K", "keys"), ("V", "values")):
                x = torch.cat([getattr(self.layers[i], attr) for i in info.layers], dim=1).transpose(1, 2).reshape(S, -1)
                g = torch.zeros(x.shape[1], x.shape[1], dtype=torch.float64, device="cuda")
                for lo in range(0, S, 8192):
                    blk = x[lo:lo + 8192].double()
                    g += blk.t() @ blk
                grams[(last_layer_idx, name)] = (g, x.clone())
        return orig_merge(self, last_layer_idx)

    FakeLayerMergingCache.grouped_layer_merging = spy
    try:
        with torch.no_grad():
            out = model.generate(ids, max_new_tokens=8, do_sample=False, return_dict_in_generate=True, output_logits=True)
    finally:
        FakeLayerMergingCache.grouped_layer_merging = orig_merge
    torch.cuda.synchronize()
    assert out.sequences.shape == (1, S + 8)
    logits = torch.stack(out.logits).float()
    assert torch.isfinite(logits).all()
    cache = out.past_key_values
    assert isinstance(cache, FakeLayerMergingCache)
    for i in range(L):
        st = cache.layers[i].group
        assert st is not None and st.factors.key.A.shape[1] == 512 and st.factors.value.A.shape[1] == 768
        assert cache.layers[i].prefill_len == S
    # stored-cache error of the sampled groups against the Eckart-Young optimum at equal rank
    for (last, name), (g, x) in grams.items():
        f = cache.layers[last].group.factors.key if name == "K" else cache.layers[last].group.factors.value
        ev = torch.linalg.eigvalsh(g).flip(0).clamp_min(0)
        opt = (ev[f.rank:].sum() / ev.sum()).sqrt().item()
        num = 0.0
        for lo in range(0, S, 8192):
            rec = f.A[lo:lo + 8192].double() @ f.Vt.double()
            num += (x[lo:lo + 8192].double() - rec).pow(2).sum().item()
        ours = (num / ev.sum().item()) ** 0.5
        print(f"group ending at layer {last}, {name}: stored error {ours:.5f}, Eckart-Young optimum {opt:.5f}, ratio {ours / opt:.4f}")
        assert ours ** 2 <= (1.01 * opt) ** 2 + 3e-3 ** 2       # 1 % + the bf16 storage floor in quadrature
    del grams
    gc.collect()
    # one more token through both decode paths over the SAME cache state: fused kernel vs materialise + SDPA
    tok = out.sequences[:, -1:]
    import copy

    def step(fused):
        for layer in model.model.layers:
            layer.self_attn.xkv_fused_decode = fused
        c2 = copy.copy(cache)
        c2.layers = [copy.copy(l) for l in cache.layers]      # shallow: the factors are shared, the tails get copied on append
        for l in c2.layers:
            if l.tail_k is not None:
                l.tail_k, l.tail_v = l.tail_k.clone(), l.tail_v.clone()
        with torch.no_grad():
            return model(input_ids=tok, past_key_values=c2, use_cache=True).logits[:, -1].float()

    lg_fused = step(True)
    lg_dense = step(False)
    dev = (lg_fused - lg_dense).abs().max().item()
    scale = lg_dense.abs().max().item()
    print(f"64K decode logits: max |fused - dense-compat| = {dev:.4f} (logit scale {scale:.3f})")
    assert dev <= 5e-2 * scale
