"""GPU parity of the tcgen05 GEMM engine against torch fp32 matmul of the same bf16 operands.

bf16 x bf16 products are exact in fp32, so with small-integer data the result is exact; with
Gaussian data the only difference from torch is fp32 summation order (tolerance 1e-4 relative to
the row/column norm product)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(rows, cols, ints, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if ints:
        return torch.randint(-3, 4, (rows, cols), device="cuda", generator=g).to(torch.bfloat16)
    return torch.randn(rows, cols, device="cuda", generator=g).bfloat16()


def _run(M, N, K, a_mn, b_mn, ints=False, terms=None, transposed=False, out_bf16=False, split_k=1, sym=False,
         phases=1):
    from xkv_b200 import ops

    terms = terms or ops.TERMS_1
    nl = max(max(t) for t in terms) + 1
    A = [_mk(K, M, ints, 10 + i) if a_mn else _mk(M, K, ints, 10 + i) for i in range(nl)]
    if sym:
        B = A
    else:
        B = [_mk(K, N, ints, 20 + i) if b_mn else _mk(N, K, ints, 20 + i) for i in range(nl)]
    ref = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    for ta, tb in terms:
        a = A[ta].float().t() if a_mn else A[ta].float()
        b = B[tb].float().t() if b_mn else B[tb].float()
        ref += a @ b.t()
    shape = (N, M) if transposed else (M, N)
    dt = torch.bfloat16 if out_bf16 else torch.float32
    out = torch.full((split_k,) + shape, float("nan"), device="cuda", dtype=dt)
    p = ops.make_problem(A, B, out[0], M=M, N=N, K=K, a_mn_major=a_mn, b_mn_major=b_mn, terms=terms,
                         out_transposed=transposed, sym_upper=sym, split_k=split_k,
                         split_stride=out.stride(0), accum_phases=phases)
    ops.gemm_grouped([p])
    torch.cuda.synchronize()
    got = out.float().sum(0)
    if transposed:
        got = got.t()
    return got, ref


def _check(got, ref, exact, sym=False, out_bf16=False):
    if sym:
        # only tiles touching the upper triangle are defined
        mask = torch.triu(torch.ones_like(ref, dtype=torch.bool))
        assert not torch.isnan(got[mask]).any()
        got = torch.where(mask, got, ref)
    assert not torch.isnan(got).any(), "unwritten output elements"
    if exact and not out_bf16:
        assert torch.equal(got, ref), f"max abs diff {(got - ref).abs().max().item()}"
    else:
        tol = 1e-2 if out_bf16 else 1e-4
        scale = ref.abs().max().item() + 1e-6
        assert (got - ref).abs().max().item() <= tol * scale


MAJORS = [(False, False), (False, True), (True, False), (True, True)]


@pytest.mark.parametrize("a_mn,b_mn", MAJORS)
def test_single_tile_exact(a_mn, b_mn):
    got, ref = _run(128, 256, 64, a_mn, b_mn, ints=True)
    _check(got, ref, exact=True)


@pytest.mark.parametrize("a_mn,b_mn", MAJORS)
def test_multi_kblock_exact(a_mn, b_mn):
    got, ref = _run(256, 512, 512, a_mn, b_mn, ints=True)
    _check(got, ref, exact=True)


@pytest.mark.parametrize("a_mn,b_mn", MAJORS)
def test_ragged_shapes(a_mn, b_mn):
    # M, N not multiples of the tile, K not a multiple of 64 (TMA zero-fills the tails)
    got, ref = _run(200, 328, 168, a_mn, b_mn, ints=True)
    _check(got, ref, exact=True)


@pytest.mark.parametrize("a_mn,b_mn", MAJORS)
def test_gaussian(a_mn, b_mn):
    got, ref = _run(384, 768, 1024, a_mn, b_mn)
    _check(got, ref, exact=False)


def test_terms_and_split_k():
    from xkv_b200 import ops

    got, ref = _run(256, 256, 1024, False, False, ints=True, terms=ops.TERMS_6, split_k=4)
    _check(got, ref, exact=True)
    got, ref = _run(130, 248, 640, False, True, ints=True, terms=ops.TERMS_3, split_k=3)
    _check(got, ref, exact=True)


def test_transposed_and_bf16_outputs():
    got, ref = _run(256, 320, 256, False, False, ints=True, transposed=True)
    _check(got, ref, exact=True)
    got, ref = _run(256, 512, 256, False, False, out_bf16=True)
    _check(got, ref, exact=False, out_bf16=True)
    got, ref = _run(200, 304, 128, False, True, out_bf16=True, transposed=True)
    _check(got, ref, exact=False, out_bf16=True)


def test_symmetric_gram_tiles():
    # G = X^T X with X (K x n) row-major: both operands MN-major views of the same matrix
    got, ref = _run(768, 768, 1024, True, True, ints=True, sym=True, split_k=2)
    _check(got, ref, exact=True, sym=True)


@pytest.mark.parametrize("phases", [2, 3, 5, 64])
def test_accumulation_phases_exact(phases):
    """The k range cut into pieces that alternate between two TMEM accumulators and are summed by the epilogue
    (the Gram over long token ranges): exact on integer data for even / odd piece counts, more pieces than k-blocks
    (K = 1100 -> 18 k-blocks), ragged tiles, the symmetric tile set and split-K on top."""
    got, ref = _run(200, 328, 1100, True, True, ints=True, phases=phases)
    _check(got, ref, exact=True)
    got, ref = _run(768, 768, 2048, True, True, ints=True, sym=True, split_k=2, phases=phases)
    _check(got, ref, exact=True, sym=True)
    from xkv_b200 import ops

    got, ref = _run(256, 256, 1024, False, False, ints=True, terms=ops.TERMS_6, phases=phases)
    _check(got, ref, exact=True)


def test_accumulation_phases_need_fp32_output():
    from xkv_b200 import _lib

    with pytest.raises(_lib.XkvError):
        _run(256, 512, 256, False, False, out_bf16=True, phases=2)


def test_run_if_predicate_skips_a_problem():
    from xkv_b200 import ops

    a, b = _mk(128, 64, True, 1), _mk(256, 64, True, 2)
    flags = torch.tensor([0, 1], device="cuda", dtype=torch.int32)
    outs = [torch.full((128, 256), float("nan"), device="cuda") for _ in range(2)]
    ps = []
    for i, o in enumerate(outs):
        p = ops.make_problem([a], [b], o, M=128, N=256, K=64)
        p.run_if = flags[i:].data_ptr()
        ps.append(p)
    ops.gemm_grouped(ps)
    torch.cuda.synchronize()
    assert torch.isnan(outs[0]).all()                       # flag 0: untouched
    assert torch.equal(outs[1], a.float() @ b.float().t())  # flag 1: computed


def test_grouped_launch():
    from xkv_b200 import ops

    torch.manual_seed(1)
    probs, refs, outs, keep = [], [], [], []
    for (M, N, K) in [(128, 256, 128), (300, 100, 256), (640, 640, 512)]:
        A = torch.randint(-3, 4, (M, K), device="cuda").to(torch.bfloat16)
        B = torch.randint(-3, 4, (N, K), device="cuda").to(torch.bfloat16)
        keep += [A, B]  # problems hold raw pointers: the operands must outlive the launch
        out = torch.full((M, N), float("nan"), device="cuda")
        probs.append(ops.make_problem([A], [B], out, M=M, N=N, K=K))
        refs.append(A.float() @ B.float().t())
        outs.append(out)
    ops.gemm_grouped(probs)
    torch.cuda.synchronize()
    for o, r in zip(outs, refs):
        assert torch.equal(o, r)


def _gram(n, K, *, layers=0, split_k=1, phases=1, ints=True, seed=5):
    """Upper-triangle tiles of G = X^T X through xkv_gemm_grouped; X (K x n) packed or as `layers` column blocks."""
    from xkv_b200 import ops

    X = _mk(K, n, ints, seed)
    out = torch.full((split_k, n, n), float("nan"), device="cuda")
    kw = dict(M=n, N=n, K=K, a_mn_major=True, b_mn_major=True, sym_upper=True, split_k=split_k,
              split_stride=out.stride(0), accum_phases=phases)
    if layers:
        cols = n // layers
        # per-layer tensors as the cache holds them: separate allocations with their own leading dimension
        parts = [X[:, i * cols:(i + 1) * cols].contiguous() for i in range(layers)]
        p = ops.make_problem([], [], out[0], a_layers=parts, b_layers=parts, **kw)
    else:
        parts = [X]
        p = ops.make_problem([X], [X], out[0], **kw)
    ops.gemm_grouped([p])
    torch.cuda.synchronize()
    return out.sum(0), X.float().t() @ X.float()


@pytest.mark.parametrize("n,K,layers,split_k,phases", [
    (256, 64, 0, 1, 1),        # one pair tile, one k block
    (768, 1024, 0, 2, 1),      # 3 x 3 pair tiles (6 in the upper set), split-K
    (1024, 4096, 4, 1, 4),     # layered operands read in place, accumulation phases
    (200, 1100, 0, 1, 3),      # ragged: the odd CTA's rows and columns are partly out of range
    (328, 168, 0, 1, 1),       # n not a multiple of 64, K not a multiple of 64; the odd CTA of tile row 1 is all padding
    (1536, 2048, 3, 3, 5),     # 3 layers of 512 columns: pair tiles straddle layer boundaries
])
def test_gram_pair_kernel_matches_single_cta_tiles_bit_for_bit(n, K, layers, split_k, phases):
    """The CTA-pair Gram (cta_group::2, 256 x 256 tiles) against torch (exact on integer data) and against the single-CTA
    tiles of the same engine on Gaussian data: the K reduction of an output element runs in the same order in both, so the
    upper-triangle tiles agree bit for bit."""
    from xkv_b200 import _lib

    lib = _lib.load()
    try:
        lib.xkv_gemm_set_gram_pair(1)
        got, ref = _gram(n, K, layers=layers, split_k=split_k, phases=phases)
        _check(got, ref, exact=True, sym=True)
        pair, _ = _gram(n, K, layers=layers, split_k=split_k, phases=phases, ints=False)
        lib.xkv_gemm_set_gram_pair(0)
        single, ref = _gram(n, K, layers=layers, split_k=split_k, phases=phases, ints=False)
    finally:
        lib.xkv_gemm_set_gram_pair(1)
    mask = torch.triu(torch.ones_like(ref, dtype=torch.bool))
    assert not torch.isnan(pair[mask]).any()
    assert torch.equal(pair[mask], single[mask])
    _check(pair, ref, exact=False, sym=True)


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,kw", [
    (4096, 576, 1024, False, False, dict(terms="T3", transposed=True)),     # power step, roles swapped: N = 3 x 192
    (2048, 832, 512, False, False, dict(terms="T6", transposed=True)),      # N = 4 x 208
    (1024, 576, 576, True, False, dict(terms="T6", transposed=True)),       # triangular solve, roles swapped
    (2304, 512, 1024, False, False, dict(out_bf16=True)),                   # projection form
    (2100, 768, 640, False, False, dict(out_bf16=True, split_k=1)),         # ragged M, 3 N tiles
    (512, 100, 256, False, False, dict()),                                  # N tile 112, mostly padding in the last 16
    (768, 40, 192, True, False, dict(transposed=True, split_k=2)),          # N tile 48
    (456, 256, 320, False, True, dict(terms="T3")),                         # MN-major B: 64-column chunks
])
def test_pair_kernels_match_single_cta_tiles(M, N, K, a_mn, b_mn, kw):
    """Every product form the CTA-pair kernel takes over (cta_group::2, 256 x bn tiles, bn a multiple of 16) against torch
    (exact on integer data) and bit for bit against the single-CTA tiles on Gaussian data."""
    from xkv_b200 import _lib, ops

    kw = dict(kw)
    terms = {"T3": ops.TERMS_3, "T6": ops.TERMS_6, None: None}[kw.pop("terms", None)]
    lib = _lib.load()
    try:
        lib.xkv_gemm_set_gram_pair(1)
        got, ref = _run(M, N, K, a_mn, b_mn, ints=True, terms=terms, **kw)
        _check(got, ref, exact=True, out_bf16=kw.get("out_bf16", False))
        pair, _ = _run(M, N, K, a_mn, b_mn, terms=terms, **kw)
        lib.xkv_gemm_set_gram_pair(0)
        single, _ = _run(M, N, K, a_mn, b_mn, terms=terms, **kw)
    finally:
        lib.xkv_gemm_set_gram_pair(1)
    assert torch.equal(pair, single)


def test_pair_kernel_layered_projection_form():
    """A = X V with X the column-wise concatenation of layer matrices read in place (K-major layered A, the k range
    crosses layer boundaries), bf16 output: the pair kernel against the packed operand on single-CTA tiles."""
    from xkv_b200 import _lib, ops

    lib = _lib.load()
    S, lc, nl, r = 1280, 256, 3, 192
    layers = [_mk(S, lc, False, 40 + i) for i in range(nl)]
    X = torch.cat(layers, dim=1).contiguous()
    V = _mk(r, nl * lc, False, 50)
    outs = []
    try:
        for pair, layered in ((1, True), (0, True), (0, False)):
            lib.xkv_gemm_set_gram_pair(pair)
            out = torch.full((S, r), float("nan"), device="cuda", dtype=torch.bfloat16)
            if layered:
                p = ops.make_problem([], [V], out, M=S, N=r, K=nl * lc, a_layers=layers)
            else:
                p = ops.make_problem([X], [V], out, M=S, N=r, K=nl * lc)
            ops.gemm_grouped([p])
            torch.cuda.synchronize()
            outs.append(out)
    finally:
        lib.xkv_gemm_set_gram_pair(1)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ref = (X.float() @ V.float().t())
    assert (outs[0].float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
