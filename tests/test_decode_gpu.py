"""GPU parity of the fused decode attention (reconstruct + RoPE + GQA softmax over the factored cache)
against the oracle: dense K^ = bf16(A Vk_l^T), HF RoPE in bf16 (cache:142-148), cat-append of the decode
tokens (cache:129) and SDPA with repeat_kv (llama.py:58-69).  Tolerance: bf16 (2e-2 of the output scale)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(S, H, D, qpk, rk, rv, T, layers_in_group, layer, rope, seed=0, dim_major_tables=True):
    from oracle import xkv_oracle as O
    from xkv_b200 import ops, synthetic

    g = torch.Generator().manual_seed(seed)
    n = layers_in_group * H * D
    Hq = H * qpk
    a_k = (torch.randn(S, rk, generator=g) * 0.6).bfloat16()
    a_v = torch.randn(S, rv, generator=g).bfloat16()
    v_k = torch.linalg.qr(torch.randn(n, rk, generator=g))[0].contiguous().bfloat16()
    v_v = torch.linalg.qr(torch.randn(n, rv, generator=g))[0].contiguous().bfloat16()
    q = torch.randn(Hq, D, generator=g).bfloat16()
    k_tail = torch.randn(H, max(T, 1), D, generator=g).bfloat16()[:, :T]
    v_tail = torch.randn(H, max(T, 1), D, generator=g).bfloat16()[:, :T]
    cos, sin = synthetic.llama3_rope(S, D)
    cos, sin = cos[0], sin[0]
    rows = slice(layer * H * D, (layer + 1) * H * D)
    # ---- oracle on the CPU ----
    k_hat = (a_k.float() @ v_k[rows].float().t()).bfloat16().view(S, H, D).permute(1, 0, 2)[None]   # (1,H,S,D)
    v_hat = (a_v.float() @ v_v[rows].float().t()).bfloat16().view(S, H, D).permute(1, 0, 2)[None]
    if rope:
        k_hat = O.apply_rope(k_hat, cos[None], sin[None])
    kt = k_tail[None] if T else None
    vt = v_tail[None] if T else None
    ref = O.decode_attention(q[None, :, None, :].float(), k_hat.float(), v_hat.float(),
                             kt.float() if T else None, vt.float() if T else None, scaling=1.0 / math.sqrt(D))[0, :, 0]
    # ---- CUDA path ----
    dev = "cuda"
    rope_t = ops.rope_tables_dim_major(cos.to(dev), sin.to(dev)) if (rope and dim_major_tables and D == 128) else None
    out = ops.decode_attention(q.to(dev), a_k.to(dev), v_k.to(dev)[rows], a_v.to(dev), v_v.to(dev)[rows], H,
                               cos.to(dev) if rope else None, sin.to(dev) if rope else None,
                               k_tail.to(dev) if T else None, v_tail.to(dev) if T else None, 1.0 / math.sqrt(D),
                               rope_t=rope_t)
    torch.cuda.synchronize()
    got = out.float().cpu()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    print(f"S={S} H={H} D={D} qpk={qpk} rk={rk} rv={rv} T={T} rope={rope}: max|diff|={err:.4f} (out scale {scale:.3f})")
    assert torch.isfinite(got).all()
    assert err <= 2e-2 * max(scale, 1e-3)


@pytest.mark.parametrize(
    "S,H,D,qpk,rk,rv,T,G,layer,rope",
    [
        (512, 2, 128, 4, 64, 128, 0, 2, 1, True),      # smallest: two token tiles... one n-tile
        (1000, 8, 128, 4, 128, 192, 7, 4, 2, True),    # ragged token count, Llama head layout, decode tail
        (4096, 8, 128, 4, 512, 768, 33, 4, 3, True),   # config-1 ranks
        (777, 4, 64, 2, 64, 64, 3, 1, 0, True),        # head_dim 64, single-layer group
        (2048, 1, 128, 8, 128, 128, 5, 4, 1, False),   # one kv head, 8 q heads, no RoPE re-application
        (300, 3, 128, 1, 96, 160, 1, 2, 0, True),      # MHA (qpk 1), 3 heads: second n-tile half empty, rank not /64
    ],
)
@pytest.mark.parametrize("variant", ["auto", "tr", "tiled", "ffma", "pair"])
def test_decode_attention_matches_oracle(S, H, D, qpk, rk, rv, T, G, layer, rope, variant):
    """auto: persistent scores kernel with the right-factor slice resident in shared memory (for head_dim 128 the rotated
    keys go back to TMEM and a second MMA contracts them with q); tr: the transposed persistent kernel (right-factor
    slice resident in TENSOR memory, token stream as the shared-memory operand, RoPE partners meet by a warp shuffle,
    dim-major RoPE tables; head_dim 128 and r_k <= 512); tiled: the tile-per-CTA kernel used when that slice
    does not fit; ffma: persistent kernel with the FFMA epilogue (head_dim 64 path); pair: cta_group::2 CTA pairs."""
    from xkv_b200 import _lib

    lib = _lib.load()
    lib.xkv_decode_force_tiled(int(variant == "tiled"))
    lib.xkv_decode_set_variant({"auto": 0, "tr": 4, "tiled": 0, "ffma": 1, "pair": 3}[variant])
    try:
        _case(S, H, D, qpk, rk, rv, T, G, layer, rope, dim_major_tables=(variant in ("auto", "tr")))
    finally:
        lib.xkv_decode_force_tiled(0)
        lib.xkv_decode_set_variant(0)


@pytest.mark.parametrize("S,T", [(300, 700), (256, 90), (1024, 3000)])
def test_dense_tail_longer_than_a_softmax_chunk(S, T):
    """Long generations after a short prompt: the dense decode tail spans several of the softmax's 16 chunks (each CTA
    scores the tail tokens of its own chunk; an earlier version let the last chunk's CTA score all of them while the other
    CTAs were already reading those scores)."""
    _case(S, 2, 128, 4, 64, 64, T, 2, 1, True)
    _case(S, 4, 64, 2, 32, 64, T, 1, 0, False)


def test_decode_large_rank_falls_back_to_tiled_kernel():
    # r_k = 1024: one head's slice is 256 KiB > 128 KiB of shared memory
    _case(1024, 2, 128, 4, 1024, 256, 2, 4, 1, True)


def test_transposed_kernel_all_k_blocks_in_tensor_memory_and_the_eighth_in_shared_memory():
    # r_k = 448: seven 64-wide K blocks, all in TMEM; r_k = 512: the eighth block goes through shared memory (SS form);
    # r_k = 456: ragged last block; several token tiles per CTA at 148 CTAs needs S > 148 * 128 / H
    from xkv_b200 import _lib

    _lib.load().xkv_decode_set_variant(4)
    try:
        _tr_cases()
    finally:
        _lib.load().xkv_decode_set_variant(0)


def _tr_cases():
    _case(3000, 8, 128, 4, 448, 192, 2, 4, 1, True)
    _case(3000, 8, 128, 4, 456, 192, 0, 4, 0, True)
    _case(40000, 8, 128, 4, 512, 256, 3, 4, 2, True)
    _case(5000, 2, 128, 8, 512, 128, 1, 2, 1, False)


def test_token_shards_merge_to_the_unsharded_attention():
    """SURVEY section 8e, decode with token shards: each shard runs the fused kernel on ITS rows of A_k / A_v with the RoPE
    rows of its positions and returns (normalised output, log-sum-exp); the flash-decoding merge must reproduce the
    kernel's own output over the whole context.  The dense decode tail lives on the last shard."""
    from xkv_b200 import ops, parallel, synthetic

    S, H, D, qpk, rk, rv, T = 5000, 8, 128, 4, 256, 384, 3
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(11)
    a_k = (torch.randn(S, rk, generator=g, device=dev) * 0.6).bfloat16()
    a_v = torch.randn(S, rv, generator=g, device=dev).bfloat16()
    v_k = torch.linalg.qr(torch.randn(H * D, rk, generator=g, device=dev))[0].contiguous().bfloat16()
    v_v = torch.linalg.qr(torch.randn(H * D, rv, generator=g, device=dev))[0].contiguous().bfloat16()
    q = torch.randn(H * qpk, D, generator=g, device=dev).bfloat16()
    k_tail = torch.randn(H, T, D, generator=g, device=dev).bfloat16()
    v_tail = torch.randn(H, T, D, generator=g, device=dev).bfloat16()
    cos, sin = synthetic.llama3_rope(S, D, device=dev)
    cos, sin = cos[0].contiguous(), sin[0].contiguous()
    scale = 1.0 / math.sqrt(D)
    lse_full = torch.empty(H * qpk, device=dev)
    full = ops.decode_attention(q, a_k, v_k, a_v, v_v, H, cos, sin, k_tail, v_tail, scale, lse_out=lse_full)
    world = 3
    outs, lses = [], []
    for rank in range(world):
        b, e = parallel.token_shard(S, world, rank)
        last = rank == world - 1
        lse = torch.empty(H * qpk, device=dev)
        o = ops.decode_attention(q, a_k[b:e], v_k, a_v[b:e], v_v, H, cos[b:e], sin[b:e], k_tail if last else None,
                                 v_tail if last else None, scale, lse_out=lse)
        outs.append(o)
        lses.append(lse)
    merged, lse = parallel.merge_partial_attention(torch.stack(outs), torch.stack(lses))
    torch.cuda.synchronize()
    scale_out = full.float().abs().max().item()
    err = (merged - full.float()).abs().max().item()
    print(f"token shards x{world}: max|diff| = {err:.5f} (output scale {scale_out:.3f}), lse diff {(lse - lse_full).abs().max().item():.2e}")
    assert err <= 2e-2 * scale_out
    assert (lse - lse_full).abs().max().item() < 1e-3


def test_rope_tables_dim_major():
    from xkv_b200 import ops, synthetic

    S, D = 1000, 128
    cos, sin = synthetic.llama3_rope(S, D)
    cos, sin = cos[0].cuda(), sin[0].cuda()
    ct, st = ops.rope_tables_dim_major(cos, sin, capacity=1100)
    torch.cuda.synchronize()
    assert ct.shape == (64, 1152) and st.shape == (64, 1152)
    assert torch.equal(ct[:, :S], cos[:, :64].t()) and torch.equal(st[:, :S], sin[:, :64].t())
    assert ct[:, S:].abs().max().item() == 0 and st[:, S:].abs().max().item() == 0


def test_rope_bf16_matches_hf_formula():
    from oracle import xkv_oracle as O
    from xkv_b200 import ops, synthetic

    torch.manual_seed(0)
    S, H, D = 257, 8, 128
    x = torch.randn(S, H, D).bfloat16()
    cos, sin = synthetic.llama3_rope(S, D)
    ref = O.apply_rope(x.permute(1, 0, 2)[None], cos, sin)[0].permute(1, 0, 2)
    got = ops.rope_bf16_(x.cuda().clone(), cos[0].cuda(), sin[0].cuda())
    torch.cuda.synchronize()
    assert torch.equal(got.cpu(), ref)
