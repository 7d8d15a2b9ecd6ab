"""Low-rank factorisation of a layer group's token-major KV matrix on the B200 kernels.

Replaces the reference's ``fake_svd`` (fake_layer_merge_dynamic_cache.py:11-29: exact
``torch.linalg.svd`` -> truncate -> multiply back) by a factorisation that is *stored*:

    X (m x n, bf16)  ~=  A (m x r, bf16) @ Vt (r x n, bf16),    A = X V,  V orthonormal columns

so that ``A @ Vt`` is the orthogonal projection of X onto (an accurate approximation of) its dominant
rank-r right singular subspace, i.e. the same matrix ``U_r S_r V_r^T`` the reference multiplies back.

Algorithm (every product is a tcgen05 GEMM of xkv_gemm.cu, everything stays on the device and on the
caller's stream; nothing synchronises):

  1. Gram            G = X^T X            one pass over X, symmetric tile set, fp32 accumulation in TMEM
  2. range finder    Y = G Omega           Omega Gaussian n x l, l = r + oversample
  3. CholeskyQR      Q = orth(Y)           row-normalise, S = Y^T Y (6-term bf16-limb product ~ fp32),
                                           blocked Cholesky + explicit inverse, Q = Y L^-T; repeated
  4. power steps     Y = G Q; CholeskyQR   `power_iters` times (3-term limb product)
  5. Rayleigh-Ritz   shared-memory Jacobi on the window of T = Q^T G Q that straddles column r: the
                     triangular (Cholesky) orthogonalisation keeps the columns ordered by dominance,
                     so only the boundary window needs rotating ("windowed Rayleigh-Ritz")
  6. right factor    V = first r columns;  Vt / V in bf16
  7. projection      A = X V               second and last pass over X

The Gram route is exact up to fp32 rounding of G, needs exactly two reads of X, and was chosen over
sketching X directly because r is a sizeable fraction of n here (512..768 of 4096).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import ctypes as C

import torch

from . import _lib, ops
from ._lib import XkvError


@dataclass
class FactorizeOptions:
    power_iters: int = 6
    oversample: int = 64
    first_passes: int = 2      # CholeskyQR passes after the range finder (ill-conditioned sketch)
    passes: int = 2            # CholeskyQR passes after each power step
    final_passes: int = 2      # passes after the last power step (orthonormal to ~1e-5 after two)
    window: int = 128          # Rayleigh-Ritz window width (<= 160: A and V live in shared memory)
    jacobi_sweeps: int = 6
    rayleigh_ritz: bool = True
    want_sigma: bool = True    # also diagonalise the leading window to report singular values
    gram_split_k: int = 1
    small_split_k: int = 8
    shifts: tuple = (3e-4, 1e-6, 1e-7)   # diagonal shift of CholeskyQR pass 0, 1, 2, ...
    pivot_floor: float = 1e-12
    seed: int = 1234
    profile: bool = False


@dataclass
class Factors:
    """Stored form of one compressed (group, K-or-V) matrix."""

    A: torch.Tensor                     # (m, r)  bf16  token-side factor  (= U_r S_r up to rotation)
    Vt: torch.Tensor                    # (r, n)  bf16  right factor, rows = basis vectors
    V: torch.Tensor                     # (n, r)  bf16  same, transposed (layout the decode kernel reads)
    rank: int
    sigma_lead: Optional[torch.Tensor] = None   # leading singular values (Ritz estimates), fp32
    A_storage: Optional[torch.Tensor] = None    # A plus spare rows for appended tokens (A is a view of it)
    timings: Dict[str, float] = field(default_factory=dict)

    def reconstruct(self) -> torch.Tensor:
        """Dense bf16 reconstruction (test/debug helper; the product path never materialises it)."""
        return (self.A.float() @ self.Vt.float()).to(torch.bfloat16)


def _round_up(x: int, k: int) -> int:
    return (x + k - 1) // k * k


def sketch_width(rank: int, oversample: int = 64) -> int:
    return _round_up(rank + oversample, 64)


def _chunks(seq: Sequence, size: int):
    for i in range(0, len(seq), size):
        yield seq[i:i + size]


def _gemm(problems: List) -> None:
    for chunk in _chunks(problems, 16):
        ops.gemm_grouped(chunk)


class _Timer:
    def __init__(self, enabled: bool):
        self.enabled = enabled
        self.events = []

    def mark(self, name: str) -> None:
        if self.enabled:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.events.append((name, e))

    def result(self) -> Dict[str, float]:
        if not self.enabled or len(self.events) < 2:
            return {}
        torch.cuda.synchronize()
        out: Dict[str, float] = {}
        for (_, e0), (name, e1) in zip(self.events[:-1], self.events[1:]):
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        return out


_STAGES = ("gram_gemm", "gram_reduce_split", "range_finder", "power_iters", "rayleigh_ritz", "project")


def _c_options(opts: FactorizeOptions) -> "_lib.FactorizeOptions":
    o = _lib.FactorizeOptions()
    _lib.load().xkv_factorize_default_options(C.byref(o))
    o.power_iters, o.oversample = opts.power_iters, opts.oversample
    o.first_passes, o.passes, o.final_passes = opts.first_passes, opts.passes, opts.final_passes
    o.window, o.jacobi_sweeps = opts.window, opts.jacobi_sweeps
    o.rayleigh_ritz, o.want_sigma = int(opts.rayleigh_ritz), int(opts.want_sigma)
    o.gram_split_k, o.small_split_k = opts.gram_split_k, opts.small_split_k
    for i in range(4):
        o.shifts[i] = opts.shifts[min(i, len(opts.shifts) - 1)]
    o.pivot_floor = opts.pivot_floor
    o.seed = opts.seed
    return o


def workspace_bytes(batch: int, m: int, n: int, rank: int, opts: Optional[FactorizeOptions] = None) -> int:
    o = _c_options(opts or FactorizeOptions())
    return int(_lib.load().xkv_factorize_workspace_bytes(batch, m, n, rank, C.byref(o)))


def factorize_batch(xs: Sequence[torch.Tensor], rank: int, opts: Optional[FactorizeOptions] = None,
                    workspace: Optional[torch.Tensor] = None, process_group=None, extra_rows: int = 0) -> List[Factors]:
    """Factorise a batch of equally-shaped token-major matrices (m x n bf16) at rank `rank`.

    One call into the library's stream-ordered driver (xkv_factorize_batch); batches larger than the
    C-ABI's per-call limit are processed in chunks that reuse one workspace.

    With `process_group` (torch.distributed, NCCL) the inputs are the LOCAL token shards of matrices whose
    rows are split over the group's ranks: the local Gram matrices are summed with one all-reduce, every
    rank derives the same right factor, and `A` holds the local rows only."""
    opts = opts or FactorizeOptions()
    if len(xs) == 0:
        return []
    lib = _lib.load()
    m, n = xs[0].shape
    dev = xs[0].device
    for x in xs:
        if not x.is_cuda:
            raise XkvError("factorize: CUDA tensors required (no CPU path)")
        if x.dtype != torch.bfloat16 or tuple(x.shape) != (m, n) or x.stride(1) != 1 or x.stride(0) != xs[0].stride(0):
            raise XkvError("factorize: inputs must be equally-shaped row-major bf16 matrices")
    r = int(rank)
    co = _c_options(opts)
    chunk = min(len(xs), _lib.MAX_BATCH)
    need = int(lib.xkv_factorize_workspace_bytes(chunk, m, n, r, C.byref(co)))
    if need == 0:
        raise XkvError(lib.xkv_last_error().decode() or "factorize: invalid problem")
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    nsig = int(lib.xkv_factorize_sigma_count(r, C.byref(co)))
    out: List[Factors] = []
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    all_events = []
    for lo in range(0, len(xs), chunk):
        part = xs[lo:lo + chunk]
        nb = len(part)
        # `extra_rows` spare token rows behind A (for tokens appended later by xkv_append_project)
        a_store = [torch.empty(m + extra_rows, r, dtype=torch.bfloat16, device=dev) for _ in range(nb)]
        a = [t[:m] for t in a_store]
        vt = [torch.empty(r, n, dtype=torch.bfloat16, device=dev) for _ in range(nb)]
        v = [torch.empty(n, r, dtype=torch.bfloat16, device=dev) for _ in range(nb)]
        sig = [torch.empty(nsig, dtype=torch.float32, device=dev) if nsig else None for _ in range(nb)]
        events = [torch.cuda.Event(enable_timing=True) for _ in range(7)] if opts.profile else None
        ev_arr = None
        if events is not None:
            for e in events:
                e.record()  # torch creates the cudaEvent lazily; recording materialises the handle
            ev_arr = (C.c_void_p * 7)(*[e.cuda_event for e in events])
            all_events.append(events)
        def call(phase, grams):
            _lib.check(lib.xkv_factorize_batch(
                ops._ptr_array(part), nb, m, n, part[0].stride(0), r, C.byref(co), ops._ptr_array(a),
                ops._ptr_array(vt), ops._ptr_array(v), ops._ptr_array(sig) if nsig else None,
                ops._ptr_array(grams) if grams is not None else None, phase, C.c_void_p(workspace.data_ptr()),
                workspace.numel() * workspace.element_size(), ev_arr, stream))

        if process_group is None:
            call(0, None)
        else:
            import torch.distributed as dist

            grams = torch.empty(nb, n, n, dtype=torch.float32, device=dev)
            call(1, list(grams))
            dist.all_reduce(grams, op=dist.ReduceOp.SUM, group=process_group)
            call(2, list(grams))
        for b in range(nb):
            out.append(Factors(A=a[b], Vt=vt[b], V=v[b], rank=r, sigma_lead=sig[b], A_storage=a_store[b]))
    if opts.profile:
        torch.cuda.synchronize()
        timings: Dict[str, float] = {}
        for events in all_events:
            for i, name in enumerate(_STAGES):
                timings[name] = timings.get(name, 0.0) + events[i].elapsed_time(events[i + 1])
        for f in out:
            f.timings = timings
    return out


def factorize_batch_py(xs: Sequence[torch.Tensor], rank: int, opts: Optional[FactorizeOptions] = None) -> List[Factors]:
    """Same pipeline orchestrated from Python, kernel by kernel (debug aid and cross-check of the C driver)."""
    opts = opts or FactorizeOptions()
    if len(xs) == 0:
        return []
    m, n = xs[0].shape
    dev = xs[0].device
    for x in xs:
        if not x.is_cuda:
            raise XkvError("factorize: CUDA tensors required (no CPU path)")
        if x.dtype != torch.bfloat16 or tuple(x.shape) != (m, n) or x.stride(1) != 1:
            raise XkvError("factorize: inputs must be equally-shaped row-major bf16 matrices")
    r = int(rank)
    l = sketch_width(r, opts.oversample)
    if r <= 0 or r > m or l > n or n % 8 != 0:
        raise XkvError(f"factorize: rank {r} (sketch {l}) does not fit a {m} x {n} matrix")
    B = len(xs)
    f32, bf16 = torch.float32, torch.bfloat16
    tm = _Timer(opts.profile)
    tm.mark("start")

    def empty(*shape, dtype=f32):
        return torch.empty(*shape, dtype=dtype, device=dev)

    # ---------------- 1. Gram matrices and their bf16 limbs ----------------
    gs = opts.gram_split_k
    g_limbs = [[empty(n, n, dtype=bf16) for _ in range(3)] for _ in range(B)]
    slabs = empty(min(B, 16), gs, n, n)
    g32 = empty(n, n)
    for base in range(0, B, 16):
        idx = list(range(base, min(base + 16, B)))
        _gemm([
            ops.make_problem([xs[b]], [xs[b]], slabs[i, 0], M=n, N=n, K=m, a_mn_major=True, b_mn_major=True,
                             sym_upper=True, split_k=gs, split_stride=slabs.stride(1))
            for i, b in enumerate(idx)
        ])
        for i, b in enumerate(idx):
            ops.reduce_slabs(slabs[i], g32, symmetrize=True)
            ops.split_bf16(g32, *g_limbs[b])
    del slabs, g32
    tm.mark("gram")

    # ---------------- buffers of the subspace iteration ----------------
    f_cur = [empty(l, n) for _ in range(B)]
    f_nxt = [empty(l, n) for _ in range(B)]
    lh = [empty(l, n, dtype=bf16) for _ in range(B)]
    lm = [empty(l, n, dtype=bf16) for _ in range(B)]
    ll = [empty(l, n, dtype=bf16) for _ in range(B)]
    sk = max(1, min(opts.small_split_k, (n + 63) // 64))
    s_slabs = empty(B, sk, l, l)
    s_mat = [empty(l, l) for _ in range(B)]
    linv = [empty(l, l) for _ in range(B)]
    linv_l = [[empty(l, l, dtype=bf16) for _ in range(3)] for _ in range(B)]

    def cholqr(npass: int) -> None:
        """f_cur <- orth(f_cur) by `npass` rounds of row-normalised CholeskyQR (rows = basis vectors)."""
        nonlocal f_cur, f_nxt
        for ipass in range(npass):
            ops.normalize_rows(f_cur, lh, lm, ll)
            # pass 0 is regularised by the large shift: the 3-term product is accurate enough there (as in the C driver)
            terms = ops.TERMS_3 if ipass == 0 else ops.TERMS_6
            _gemm([
                ops.make_problem([lh[b], lm[b], ll[b]], [lh[b], lm[b], ll[b]], s_slabs[b, 0], M=l, N=l, K=n,
                                 terms=terms, sym_upper=True, split_k=sk, split_stride=s_slabs.stride(1))
                for b in range(B)
            ])
            for b in range(B):
                ops.reduce_slabs(s_slabs[b], s_mat[b], symmetrize=True)
            for chunk in _chunks(list(range(B)), 16):
                ops.cholesky_inverse([s_mat[b] for b in chunk], [linv[b] for b in chunk],
                                     opts.shifts[min(ipass, len(opts.shifts) - 1)], opts.pivot_floor)
            for b in range(B):
                ops.split_bf16(linv[b], *linv_l[b])
            # Qt = Linv * Ys   (A: Linv limbs K-major;  B: Ys limbs as [K = l][N = n], MN-major)
            _gemm([
                ops.make_problem(linv_l[b], [lh[b], lm[b], ll[b]], f_nxt[b], M=l, N=n, K=l, b_mn_major=True,
                                 terms=terms)
                for b in range(B)
            ])
            f_cur, f_nxt = f_nxt, f_cur

    def apply_gram(terms) -> None:
        """f_cur <- f_cur G  (rows stay basis vectors; G symmetric so G serves as its own K-major B operand)."""
        nonlocal f_cur, f_nxt
        nl = max(max(t) for t in terms) + 1
        _gemm([
            ops.make_problem([lh[b], lm[b], ll[b]][:max(nl, 1)], g_limbs[b][:max(nl, 1)], f_nxt[b], M=l, N=n, K=n,
                             terms=terms)
            for b in range(B)
        ])
        f_cur, f_nxt = f_nxt, f_cur

    # ---------------- 2-3. range finder ----------------
    for b in range(B):
        ops.fill_gaussian_bf16(lh[b], opts.seed + 7919 * b)
    apply_gram(ops.TERMS_1)
    cholqr(opts.first_passes)
    tm.mark("range_finder")

    # ---------------- 4. power steps ----------------
    for it in range(opts.power_iters):
        for b in range(B):
            ops.split_bf16(f_cur[b], lh[b], lm[b], None)
        apply_gram(ops.TERMS_3)
        cholqr(opts.final_passes if it == opts.power_iters - 1 else opts.passes)
    tm.mark("power_iters")

    # ---------------- 5. windowed Rayleigh-Ritz ----------------
    sigma: List[Optional[torch.Tensor]] = [None] * B
    wr = l - r
    W = min(opts.window, l)
    W -= W % 2
    wl = min(r, W - wr)
    if opts.rayleigh_ritz and wl > 0 and W >= 2 and W <= 160:
        W = wl + wr
        if W % 2:
            wl -= 1
            W -= 1
        r0 = r - wl
        wins = [(r0, True)]
        if opts.want_sigma:
            wins.append((0, False))
        nw = len(wins)
        for b in range(B):
            ops.split_bf16(f_cur[b], lh[b], lm[b], ll[b])
        yw = empty(B, nw, W, n)
        yw_l = [empty(B, nw, W, n, dtype=bf16) for _ in range(3)]
        skw = max(1, min(16, (n + 63) // 64))
        t_slabs = empty(B, nw, skw, W, W)
        t_mat = empty(B, nw, W, W)
        evals = empty(B, nw, W)
        wt = empty(B, W, W)
        # Yw = Qw G
        _gemm([
            ops.make_problem([lh[b][w0:w0 + W], lm[b][w0:w0 + W], ll[b][w0:w0 + W]], g_limbs[b], yw[b, i], M=W, N=n,
                             K=n, terms=ops.TERMS_6)
            for b in range(B) for i, (w0, _) in enumerate(wins)
        ])
        for b in range(B):
            for i in range(nw):
                ops.split_bf16(yw[b, i], yw_l[0][b, i], yw_l[1][b, i], yw_l[2][b, i])
        # Tw = Qw Yw^T
        _gemm([
            ops.make_problem([lh[b][w0:w0 + W], lm[b][w0:w0 + W], ll[b][w0:w0 + W]],
                             [yw_l[0][b, i], yw_l[1][b, i], yw_l[2][b, i]], t_slabs[b, i, 0], M=W, N=W, K=n,
                             terms=ops.TERMS_6, split_k=skw, split_stride=t_slabs.stride(2))
            for b in range(B) for i, (w0, _) in enumerate(wins)
        ])
        for b in range(B):
            for i in range(nw):
                ops.reduce_slabs(t_slabs[b, i], t_mat[b, i], symmetrize=False)
        jobs_t, jobs_e, jobs_w = [], [], []
        for b in range(B):
            for i, (_, vec) in enumerate(wins):
                jobs_t.append(t_mat[b, i])
                jobs_e.append(evals[b, i])
                jobs_w.append(wt[b] if vec else None)
        for lo in range(0, len(jobs_t), 32):
            ops.jacobi_eigh(jobs_t[lo:lo + 32], jobs_e[lo:lo + 32], jobs_w[lo:lo + 32], sweeps=opts.jacobi_sweeps)
        # rows [r0, r) of Q <- top-wl Ritz vectors of the window:  Vw = Wsel Qw
        wsel_l = [empty(B, wl, W, dtype=bf16) for _ in range(3)]
        for b in range(B):
            ops.split_bf16(wt[b][:wl], wsel_l[0][b], wsel_l[1][b], wsel_l[2][b])
        _gemm([
            ops.make_problem([wsel_l[0][b], wsel_l[1][b], wsel_l[2][b]],
                             [lh[b][r0:r0 + W], lm[b][r0:r0 + W], ll[b][r0:r0 + W]], f_cur[b][r0:r], M=wl, N=n, K=W,
                             b_mn_major=True, terms=ops.TERMS_6)
            for b in range(B)
        ])
        if opts.want_sigma:
            for b in range(B):
                sigma[b] = evals[b, 1].clamp_min(0).sqrt()
    tm.mark("rayleigh_ritz")

    # ---------------- 6. right factor in bf16 (both layouts) ----------------
    vts = [empty(r, n, dtype=bf16) for _ in range(B)]
    vs = [empty(n, r, dtype=bf16) for _ in range(B)]
    for b in range(B):
        ops.convert_bf16(f_cur[b][:r], vts[b], vs[b])

    # ---------------- 7. projection A = X V ----------------
    a_out = [empty(m, r, dtype=bf16) for _ in range(B)]
    _gemm([ops.make_problem([xs[b]], [vts[b]], a_out[b], M=m, N=r, K=n) for b in range(B)])
    tm.mark("project")
    timings = tm.result()
    return [Factors(A=a_out[b], Vt=vts[b], V=vs[b], rank=r, sigma_lead=sigma[b], timings=timings) for b in range(B)]
