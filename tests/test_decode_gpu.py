"""GPU parity of the fused decode attention (reconstruct + RoPE + GQA softmax over the factored cache)
against the oracle: dense K^ = bf16(A Vk_l^T), HF RoPE in bf16 (cache:142-148), cat-append of the decode
tokens (cache:129) and SDPA with repeat_kv (llama.py:58-69).  Tolerance: bf16 (2e-2 of the output scale)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


VARIANTS = {          # name -> (force_tiled, variant)
    "auto": (0, 0),     # head_dim 128: CTA pairs (cta_group::2), half a right-factor slice per CTA
    "pair": (0, 3),     # the same, requested explicitly
    "cl1": (0, 2),         # score MMA, one independent CTA per kv head
    "tiled": (1, 0),    # tile-per-CTA kernel (right-factor slice does not fit in shared memory)
    "ffma": (0, 1),     # persistent kernel with the FFMA epilogue (the head_dim 64 path)
    "split_rc": (0, 4), # slab reduction and combine as two launches (default: one cluster launch)
}


def _set_variant(name):
    from xkv_b200 import _lib

    lib = _lib.load()
    tiled, variant = VARIANTS[name]
    lib.xkv_decode_force_tiled(tiled)
    lib.xkv_decode_set_variant(variant)


def _case(S, H, D, qpk, rk, rv, T, layers_in_group, layer, rope, seed=0, shape="flat", device_oracle=False):
    """One fused decode call against the oracle's arithmetic.  `shape` picks the softmax the inputs produce:
      flat      q ~ N(0,1), keys of norm ~0.2: an almost uniform average (round 1's only case)
      sharp     q x 16: a handful of tokens carry the weight, chunk maxima differ by tens of units
      spike     one prefix token's key x 60: a single dominant token for the heads it aligns with (+/- 30 in score)
      tail      one DENSE tail token's key x 12: the dominant token lives in the tail chunk
      equal     A_k = 0: every prefix score is exactly 0 (all-equal scores), only the tail differs
    Checked: the output (2e-2 of its scale, bf16 tolerance) and the log-sum-exp of the scaled scores (absolute 2e-2:
    one bf16 ulp of a score of magnitude ~4; relative 1e-2 beyond that)."""
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
    if shape == "sharp":
        q = (q.float() * 16).bfloat16()
    elif shape == "spike":
        a_k[(7 * S) // 11] = (a_k[(7 * S) // 11].float() * 60).bfloat16()
    elif shape == "tail":
        assert T > 0
        k_tail[:, T // 2] = (k_tail[:, T // 2].float() * 12).bfloat16()
    elif shape == "equal":
        a_k.zero_()
    cos, sin = synthetic.llama3_rope(S, D)
    cos, sin = cos[0], sin[0]
    rows = slice(layer * H * D, (layer + 1) * H * D)
    scaling = 1.0 / math.sqrt(D)
    dev = "cuda"
    od = dev if device_oracle else "cpu"      # the oracle's formulas, on the device for the 64K cases (CPU: minutes)
    # ---- oracle arithmetic: dense K^ = bf16(A Vk_l^T), HF RoPE in bf16, cat-append of the tail, SDPA with repeat_kv ----
    k_hat = (a_k.to(od).float() @ v_k[rows].to(od).float().t()).bfloat16().view(S, H, D).permute(1, 0, 2)[None]
    v_hat = (a_v.to(od).float() @ v_v[rows].to(od).float().t()).bfloat16().view(S, H, D).permute(1, 0, 2)[None]
    if rope:
        k_hat = O.apply_rope(k_hat, cos[None].to(od), sin[None].to(od))
    kt = k_tail[None].to(od).float() if T else None
    vt = v_tail[None].to(od).float() if T else None
    ref = O.decode_attention(q.to(od)[None, :, None, :].float(), k_hat.float(), v_hat.float(), kt, vt,
                             scaling=scaling)[0, :, 0].cpu()
    k_all = torch.cat([k_hat.float(), kt], dim=-2) if T else k_hat.float()
    sc = torch.einsum("hd,hsd->hs", q.to(od).float(), k_all[0].repeat_interleave(qpk, dim=0)) * scaling
    lse_ref = torch.logsumexp(sc, dim=-1).cpu()
    # ---- CUDA path ----
    lse = torch.empty(Hq, dtype=torch.float32, device=dev)
    out = ops.decode_attention(q.to(dev), a_k.to(dev), v_k.to(dev)[rows], a_v.to(dev), v_v.to(dev)[rows], H,
                               cos.to(dev) if rope else None, sin.to(dev) if rope else None,
                               k_tail.to(dev) if T else None, v_tail.to(dev) if T else None, scaling, lse_out=lse)
    torch.cuda.synchronize()
    got = out.float().cpu()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    lse_err = ((lse.cpu() - lse_ref).abs() / lse_ref.abs().clamp(min=2.0)).max().item()
    peak = torch.softmax(sc, dim=-1).max().item()
    print(f"S={S} H={H} D={D} qpk={qpk} rk={rk} rv={rv} T={T} rope={rope} {shape}: max|diff|={err:.4f} "
          f"(out scale {scale:.3f}), lse rel diff {lse_err:.2e}, largest softmax weight {peak:.3f}")
    assert torch.isfinite(got).all() and torch.isfinite(lse).all()
    assert err <= 2e-2 * max(scale, 1e-3)
    assert lse_err <= 1e-2


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
@pytest.mark.parametrize("variant", list(VARIANTS))
def test_decode_attention_matches_oracle(S, H, D, qpk, rk, rv, T, G, layer, rope, variant):
    """Every scores kernel (and the two-launch slab reduction) against the oracle."""
    _set_variant(variant)
    try:
        _case(S, H, D, qpk, rk, rv, T, G, layer, rope)
    finally:
        _set_variant("auto")


@pytest.mark.parametrize("shape", ["sharp", "spike", "tail", "equal"])
@pytest.mark.parametrize("variant", ["auto", "cl1", "tiled", "ffma", "split_rc"])
@pytest.mark.parametrize("S,T", [(4096, 33), (4000, 100), (70, 5)])
def test_peaked_softmax_matches_oracle(S, T, variant, shape):
    """Softmax shapes a near-uniform case cannot catch: the chunk-local maxima (one chunk per split-K slab of P A_v plus
    the dense tail) differ by tens of units, so the exp(m_c - m) rescale of the slab reduction and of the combine does
    real work; S = 4000 is not a multiple of 64 and its 100-token tail spans several chunk-sized pieces; S = 70 is a
    single ragged tile."""
    _set_variant(variant)
    try:
        _case(S, 8, 128, 4, 128, 192, T, 4, 2, True, seed=3, shape=shape)
    finally:
        _set_variant("auto")


@pytest.mark.parametrize("shape", ["flat", "sharp", "spike", "tail"])
def test_peaked_softmax_at_64k_context(shape):
    """Config 2's decode shape (65536 tokens + a ragged extra 37, rank 512 / 768, 32 / 8 heads x 128) with every softmax
    shape; oracle formulas evaluated on the device."""
    _case(65536 + 37, 8, 128, 4, 512, 768, 77, 4, 1, True, seed=5, shape=shape, device_oracle=True)


@pytest.mark.parametrize("S,T", [(300, 700), (256, 90), (1024, 3000)])
def test_dense_tail_longer_than_a_softmax_chunk(S, T):
    """Long generations after a short prompt: the dense decode tail spans several of the softmax's 16 chunks (each CTA
    scores the tail tokens of its own chunk; an earlier version let the last chunk's CTA score all of them while the other
    CTAs were already reading those scores)."""
    _case(S, 2, 128, 4, 64, 64, T, 2, 1, True)
    _case(S, 4, 64, 2, 32, 64, T, 1, 0, False)


@pytest.mark.parametrize("rk,H,variant", [(768, 2, "auto"), (1024, 2, "auto"), (1024, 2, "cl1"), (1536, 4, "auto")])
def test_decode_large_ranks(rk, H, variant):
    """r_k = 768 / 1024: half a head's slice (96 / 128 KiB) still fits beside a 7- / 5-slot ring in the CTA-pair kernel;
    the single-CTA kernels cannot hold the whole slice (192 / 256 KiB > 128 KiB) and take the tile-per-CTA kernel, as every
    kernel does from r_k = 1408 up."""
    _set_variant(variant)
    try:
        _case(1024, H, 128, 4, rk, 256, 2, 4, 1, True)
    finally:
        _set_variant("auto")


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


@pytest.mark.parametrize("S,Hq,r,dr,sharp", [(600, 16, 256, 64, False), (5000, 16, 512, 64, True), (32768, 16, 512, 64, False),
                                              (70, 4, 64, 0, False), (4097, 128, 128, 8, True)])
def test_absorbed_attention_matches_dense_formula(S, Hq, r, dr, sharp):
    """xkv_decode_absorbed (the MLA latent path): s = scale (row_scale (q_hat . a_t) + bias_q . bias_k[t]),
    u = sum_t softmax(s) row_scale[t] a_t, lse — against the same formulas in fp64 on the device."""
    from xkv_b200 import ops

    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(S + r)
    a = (torch.randn(S, r, generator=g, device=dev) * 0.5).bfloat16()
    q_hat = (torch.randn(Hq, r, generator=g, device=dev) * (4.0 if sharp else 0.5)).bfloat16()
    row_scale = (0.5 + torch.rand(S, generator=g, device=dev)).float()
    bias_q = torch.randn(Hq, dr, generator=g, device=dev).bfloat16() if dr else None
    bias_k = torch.randn(S, dr, generator=g, device=dev).bfloat16() if dr else None
    scale = 0.07
    u, lse = ops.decode_absorbed(q_hat, a, scale, row_scale=row_scale, bias_q=bias_q, bias_k=bias_k)
    torch.cuda.synchronize()
    s = (q_hat.double() @ a.double().t()) * row_scale.double()[None]
    if dr:
        s = s + bias_q.double() @ bias_k.double().t()
    s = s * scale
    p = torch.softmax(s, dim=-1)
    u_ref = (p * row_scale.double()[None]) @ a.double()
    lse_ref = torch.logsumexp(s, dim=-1)
    err = (u.double() - u_ref).abs().max().item()
    sc = u_ref.abs().max().item()
    lse_err = (lse.double() - lse_ref).abs().max().item()
    print(f"absorbed S={S} Hq={Hq} r={r}: max|du| = {err:.2e} (scale {sc:.3f}), max|dlse| = {lse_err:.2e}, "
          f"largest weight {p.max().item():.3f}")
    assert torch.isfinite(u).all() and torch.isfinite(lse).all()
    assert err <= 1e-2 * max(sc, 1e-3)       # probabilities are rounded to bf16 for the tensor-core product
    assert lse_err <= 2e-3
