"""GPU: FakeLayerMergingCache on the B200 kernels against the oracle's dense cache (reference semantics:
fake_layer_merge_dynamic_cache.py:127-208) — prefill compression per group, RoPE after reconstruction,
decode tokens appended uncompressed, fused decode attention."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(G=4, H=2, S=640, D=64, rank_k=64, rank_v=128, layers=8, merge_value=True, seed=0):
    from xkv_b200 import synthetic
    from xkv_b200.configurations import generate_consecutive_xKV_config

    cfg = generate_consecutive_xKV_config(num_layers=layers, end_layer=-1, group_size=G, rank_k=rank_k, rank_v=rank_v,
                                          merge_value=merge_value)
    keys, vals = [], []
    for g in range(layers // G):
        keys += synthetic.make_group_kv(G, H, S, D, 1.0, seed + g, device="cuda")
        vals += synthetic.make_group_kv(G, H, S, D, 0.7, seed + 100 + g, device="cuda")
    cos, sin = synthetic.llama3_rope(S, D, device="cuda")
    return cfg, keys, vals, cos, sin


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("re_apply_rope", [True, False])
def test_prefill_matches_oracle_cache(re_apply_rope):
    from oracle import xkv_oracle as O
    from tests.oracle_cache import OracleCache
    from xkv_b200.customized_cache import FakeLayerMergingCache

    cfg, keys, vals, cos, sin = _setup()
    ours, ref = FakeLayerMergingCache(cfg), OracleCache(cfg)
    for l, (k, v) in enumerate(zip(keys, vals)):
        ours.update(k, v, l, mode="prefill", cos=cos, sin=sin, re_apply_rope=re_apply_rope)
        ref.update(k, v, l, mode="prefill", cos=cos, sin=sin, re_apply_rope=re_apply_rope)
    torch.cuda.synchronize()
    assert ours.get_seq_length() == ref.get_seq_length() == keys[0].shape[-2]
    for l in range(len(keys)):
        k_o, v_o = ours.materialize(l)
        k_r, v_r = ref.layers[l].keys, ref.layers[l].values
        k_true = O.apply_rope(keys[l], cos, sin) if re_apply_rope else keys[l]
        e_ours, e_ref = _rel(k_o, k_true), _rel(k_r, k_true)
        ev_ours, ev_ref = _rel(v_o, vals[l]), _rel(v_r, vals[l])
        print(f"layer {l}: K err ours {e_ours:.5f} ref {e_ref:.5f}   V err ours {ev_ours:.5f} ref {ev_ref:.5f}")
        # per-layer errors fluctuate around the group's (only the group total is optimal): 3% slack per layer
        assert e_ours <= 1.03 * e_ref and ev_ours <= 1.03 * ev_ref
        assert k_o.shape == k_r.shape and k_o.dtype == torch.bfloat16


def test_group_total_error_within_one_percent():
    from tests.oracle_cache import OracleCache
    from xkv_b200.customized_cache import FakeLayerMergingCache

    cfg, keys, vals, cos, sin = _setup()
    ours, ref = FakeLayerMergingCache(cfg), OracleCache(cfg)
    for l, (k, v) in enumerate(zip(keys, vals)):
        ours.update(k, v, l, mode="prefill", cos=cos, sin=sin, re_apply_rope=False)
        ref.update(k, v, l, mode="prefill", cos=cos, sin=sin, re_apply_rope=False)
    for g0 in (0, 4):
        for name, src, pick in (("K", keys, 0), ("V", vals, 1)):
            x = torch.cat(src[g0:g0 + 4], dim=1)
            xo = torch.cat([ours.materialize(l)[pick] for l in range(g0, g0 + 4)], dim=1)
            xr = torch.cat([(ref.layers[l].keys, ref.layers[l].values)[pick] for l in range(g0, g0 + 4)], dim=1)
            eo, er = _rel(xo, x), _rel(xr, x)
            print(f"group {g0 // 4} {name}: ours {eo:.5f} ref {er:.5f} ratio {eo / er:.4f}")
            assert eo <= 1.01 * er


def test_decode_step_matches_oracle_attention():
    from oracle import xkv_oracle as O
    from tests.oracle_cache import OracleCache
    from xkv_b200.customized_cache import FakeLayerMergingCache

    cfg, keys, vals, cos, sin = _setup()
    H, D, qpk = 2, 64, 4
    ours, ref = FakeLayerMergingCache(cfg), OracleCache(cfg)
    for l, (k, v) in enumerate(zip(keys, vals)):
        ours.update(k, v, l, mode="prefill", cos=cos, sin=sin, return_dense=False)
        ref.update(k, v, l, mode="prefill", cos=cos, sin=sin)
    g = torch.Generator(device="cuda").manual_seed(5)
    worst = 0.0
    for step in range(3):
        for l in range(len(keys)):
            q = torch.randn(1, H * qpk, 1, D, device="cuda", generator=g).bfloat16()
            kn = torch.randn(1, H, 1, D, device="cuda", generator=g).bfloat16()
            vn = torch.randn(1, H, 1, D, device="cuda", generator=g).bfloat16()
            out = ours.attend(q, kn, vn, l, 1.0 / math.sqrt(D))
            k_all, v_all = ref.update(kn, vn, l, mode="decode")
            exp = O.decode_attention(q.float(), k_all.float(), v_all.float(), scaling=1.0 / math.sqrt(D))
            assert out is not None and out.shape == exp.shape
            err = (out.float() - exp).abs().max().item() / exp.abs().max().item()
            worst = max(worst, err)
    torch.cuda.synchronize()
    print(f"decode attention vs oracle dense-cache SDPA: worst max-abs deviation {worst:.4f} of output scale")
    assert ours.get_seq_length() == keys[0].shape[-2] + 3
    assert worst < 3e-2     # bf16 tolerance (the two caches also differ by their ~0.3% subspace difference)
    # the compatibility path (dense return of update(mode='decode')) agrees with the fused kernel's inputs
    k_d, v_d = ours.materialize(0)
    assert k_d.shape[-2] == keys[0].shape[-2] + 3 and v_d.shape == k_d.shape


def test_mla_latent_slot_semantics():
    """deepseek_v2.py:217-232: one 'head' of kv_lora_rank in the key slot, k_pe in the value slot,
    re_apply_rope=False, merge_value=False (the value slot stays dense and exact)."""
    from xkv_b200 import synthetic
    from xkv_b200.configurations import generate_consecutive_xKV_config
    from xkv_b200.customized_cache import FakeLayerMergingCache

    S, R = 768, 512
    cfg = generate_consecutive_xKV_config(num_layers=3, end_layer=-1, group_size=3, rank_k=256, rank_v=None,
                                          merge_value=False)
    lat = synthetic.make_group_kv(3, 1, S, R, 1.0, 3, device="cuda")
    kpe = [torch.randn(1, 1, S, 64, device="cuda").bfloat16() for _ in range(3)]
    cache = FakeLayerMergingCache(cfg)
    outs = [cache.update(lat[l], kpe[l], l, mode="prefill", cos=None, sin=None, re_apply_rope=False) for l in range(3)]
    torch.cuda.synchronize()
    assert not cache.is_value_merged() and cache.is_key_merged()
    # earlier layers of the group see their exact latents, the last one sees the compressed ones (§3 D)
    assert torch.equal(outs[0][0], lat[0]) and not torch.equal(outs[2][0], lat[2])
    for l in range(3):
        k, v = cache.materialize(l)
        assert torch.equal(v, kpe[l])
        assert k.shape == lat[l].shape and _rel(k, lat[l]) < 0.2


def test_ungrouped_layers_and_full_rank_groups_stay_exact():
    from oracle import xkv_oracle as O
    from xkv_b200 import synthetic
    from xkv_b200.configurations import LayerGroup, xKVConfig
    from xkv_b200.customized_cache import FakeLayerMergingCache

    S, H, D = 96, 2, 64
    cfg = xKVConfig(num_layers=3, rank_k=4096, rank_v=4096, layer_groups=[LayerGroup(layers=[1, 2])])
    keys = synthetic.make_group_kv(3, H, S, D, 1.0, 1, device="cuda")
    vals = synthetic.make_group_kv(3, H, S, D, 1.0, 2, device="cuda")
    cos, sin = synthetic.llama3_rope(S, D, device="cuda")
    cache = FakeLayerMergingCache(cfg)
    for l in range(3):
        cache.update(keys[l], vals[l], l, mode="prefill", cos=cos, sin=sin)
    torch.cuda.synchronize()
    for l in range(3):   # layer 0 is un-grouped; group [1,2] has rank >= min(m, n): a no-op in the reference
        k, v = cache.materialize(l)
        k_ref = O.apply_rope(keys[l], cos, sin)
        assert torch.equal(v, vals[l]), f"layer {l}: values changed"
        assert k.shape == k_ref.shape and k.dtype == k_ref.dtype
        assert torch.equal(k, k_ref), f"layer {l}: keys differ by {(k.float() - k_ref.float()).abs().max().item()}"


def test_folding_decode_tokens_into_the_factors():
    """Extension (north-star step 4, off by default): decode tokens are projected onto the group basis once
    every layer of the group has seen them. Attention must stay close to the exact-tail result."""
    from oracle import xkv_oracle as O
    from tests.oracle_cache import OracleCache
    from xkv_b200 import synthetic
    from xkv_b200.customized_cache import FakeLayerMergingCache

    G, H, S, D, qpk, steps = 4, 2, 640, 64, 4, 5
    cfg, keys, vals, cos, sin = _setup(G=G, H=H, S=S + steps, D=D, layers=4)
    cos_all, sin_all = synthetic.llama3_rope(S + steps, D, device="cuda")
    ours = FakeLayerMergingCache(cfg, compress_decode_tokens=True, decode_capacity=64)
    ref = OracleCache(cfg)
    for l in range(4):
        ours.update(keys[l][:, :, :S], vals[l][:, :, :S], l, mode="prefill", cos=cos_all[:, :S], sin=sin_all[:, :S],
                    return_dense=False)
        ref.update(keys[l][:, :, :S], vals[l][:, :, :S], l, mode="prefill", cos=cos_all[:, :S], sin=sin_all[:, :S])
    g = torch.Generator(device="cuda").manual_seed(11)
    worst = 0.0
    for t in range(steps):
        c, s_ = cos_all[:, S + t: S + t + 1], sin_all[:, S + t: S + t + 1]
        for l in range(4):
            q = torch.randn(1, H * qpk, 1, D, device="cuda", generator=g).bfloat16()
            k_pre = keys[l][:, :, S + t: S + t + 1]          # decode tokens drawn from the same group model
            v_new = vals[l][:, :, S + t: S + t + 1]
            k_post = O.apply_rope(k_pre, c, s_)
            out = ours.attend(q, k_post, v_new, l, 1.0 / math.sqrt(D), key_pre_rope=k_pre, cos=c, sin=s_)
            k_all, v_all = ref.update(k_post, v_new, l, mode="decode")
            exp = O.decode_attention(q.float(), k_all.float(), v_all.float(), scaling=1.0 / math.sqrt(D))
            worst = max(worst, (out.float() - exp).abs().max().item() / exp.abs().max().item())
    torch.cuda.synchronize()
    st = ours._layer(0).group
    assert st.length == S + steps and all(ours._layer(l).tail_len == 0 for l in range(4))
    assert ours.get_seq_length() == S + steps
    print(f"folded decode tokens: worst attention deviation {worst:.4f} of output scale")
    assert worst < 5e-2


def test_slerp_branch_matches_oracle_cache():
    """layer_merge_impl='slerp' (SURVEY.md §8 f3): MiniCache merge of 2-layer groups, RoPE afterwards, dense
    storage; compared with the oracle cache running the reference's own formulae."""
    from oracle import xkv_oracle as O
    from xkv_b200 import synthetic
    from xkv_b200.configurations import generate_consecutive_xKV_config
    from xkv_b200.customized_cache import FakeLayerMergingCache

    H, S, D = 2, 256, 64
    cfg = generate_consecutive_xKV_config(layer_merge_impl="slerp", num_layers=2, end_layer=-1, group_size=2,
                                          slerp_t=0.5, slerp_gamma=0.05)
    keys = synthetic.make_group_kv(2, H, S, D, 1.0, 21, device="cuda")
    vals = synthetic.make_group_kv(2, H, S, D, 0.7, 22, device="cuda")
    cos, sin = synthetic.llama3_rope(S, D, device="cuda")
    cache = FakeLayerMergingCache(cfg)
    for l in range(2):
        cache.update(keys[l], vals[l], l, mode="prefill", cos=cos, sin=sin)
    torch.cuda.synchronize()
    flat = lambda t: t.float().reshape(-1, D)
    k1, k2 = O.fake_minicache_merge(flat(keys[0]), flat(keys[1]), t=0.5, gamma=0.05)
    v1, v2 = O.fake_minicache_merge(flat(vals[0]), flat(vals[1]), t=0.5, gamma=0.05)
    for l, (kr, vr) in enumerate(((k1, v1), (k2, v2))):
        k, v = cache.materialize(l)
        k_ref = O.apply_rope(kr.reshape(1, H, S, D).bfloat16(), cos, sin)
        assert (k.float() - k_ref.float()).abs().max().item() <= 3e-2 * k_ref.float().abs().max().item()
        assert (v.float() - vr.reshape(1, H, S, D)).abs().max().item() <= 2e-2 * vr.abs().max().item()
    # group size 3 is rejected like the reference does (cache:184)
    bad = generate_consecutive_xKV_config(layer_merge_impl="slerp", num_layers=3, end_layer=-1, group_size=3)
    c2 = FakeLayerMergingCache(bad)
    k3 = synthetic.make_group_kv(3, H, S, D, 1.0, 23, device="cuda")
    with pytest.raises(AssertionError, match="group size 2"):
        for l in range(3):
            c2.update(k3[l], k3[l], l, mode="prefill", cos=cos, sin=sin)
