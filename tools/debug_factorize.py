"""Step-by-step check of the factorisation pipeline against torch fp64 (debug aid, GPU only)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xkv_b200 import ops, synthetic
from xkv_b200.factorize import sketch_width


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def nans(t):
    return int(torch.isnan(t.float()).sum().item())


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    r = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    if len(sys.argv) > 4 and sys.argv[4] == "colscale":   # unmixed columns with decaying scales: near-diagonal Gram
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randn(m, n, device="cuda", generator=g)
        x *= (torch.arange(1, n + 1, device="cuda") ** -0.7)
        x = x.to(torch.bfloat16)
    elif len(sys.argv) > 4 and sys.argv[4].startswith("outliers"):   # a few channels with a huge gain
        gain = float(sys.argv[4][len("outliers"):] or 300)
        x = synthetic.group_matrix(m, n, 1.0, seed=3, device="cuda").float()
        x[:, torch.arange(8) * 97 % n] *= gain
        x = (x * (4.0 / x.abs().max())).to(torch.bfloat16)
    else:
        x = synthetic.group_matrix(m, n, 1.0, seed=1234, device="cuda")
    l = sketch_width(r)
    dev = "cuda"
    f32, bf = torch.float32, torch.bfloat16
    # 1. gram
    slabs = torch.full((1, n, n), float("nan"), device=dev)
    ops.gemm_grouped([ops.make_problem([x], [x], slabs[0], M=n, N=n, K=m, a_mn_major=True, b_mn_major=True,
                                       sym_upper=True, split_k=1, split_stride=slabs.stride(0))])
    g32 = torch.full((n, n), float("nan"), device=dev)
    ops.reduce_slabs(slabs, g32, symmetrize=True)
    torch.cuda.synchronize()
    g64 = x.double().t() @ x.double()
    print("gram: nans", nans(g32), "rel", rel(g32, g64), "sym", (g32 - g32.t()).abs().max().item())
    gl = [torch.empty(n, n, device=dev, dtype=bf) for _ in range(3)]
    ops.split_bf16(g32, *gl)
    torch.cuda.synchronize()
    print("gram limbs: rel", rel(gl[0].double() + gl[1].double() + gl[2].double(), g32))
    # 2. range finder
    lh, lm, ll = (torch.empty(l, n, device=dev, dtype=bf) for _ in range(3))
    ops.fill_gaussian_bf16(lh, 1234)
    y = torch.full((l, n), float("nan"), device=dev)
    ops.gemm_grouped([ops.make_problem([lh], [gl[0]], y, M=l, N=n, K=n)])
    torch.cuda.synchronize()
    print("Y0: nans", nans(y), "rel", rel(y, lh.double() @ gl[0].double().t()))
    cur = y
    for p in range(3):
        ops.normalize_rows([cur], [lh], [lm], [ll])
        torch.cuda.synchronize()
        ys = lh.double() + lm.double() + ll.double()
        print(f" pass {p}: normalized nans", nans(cur), "row norm dev", (cur.double().norm(dim=1) - 1).abs().max().item(),
              "limb rel", rel(ys, cur))
        sk = 8
        s_slabs = torch.full((sk, l, l), float("nan"), device=dev)
        ops.gemm_grouped([ops.make_problem([lh, lm, ll], [lh, lm, ll], s_slabs[0], M=l, N=l, K=n, terms=ops.TERMS_6,
                                           sym_upper=True, split_k=sk, split_stride=s_slabs.stride(0))])
        s = torch.full((l, l), float("nan"), device=dev)
        ops.reduce_slabs(s_slabs, s, symmetrize=True)
        torch.cuda.synchronize()
        s64 = ys @ ys.t()
        ev = torch.linalg.eigvalsh(s64)
        print(f" pass {p}: S nans", nans(s), "rel", rel(s, s64), "cond(S)", (ev[-1] / ev[0].abs()).item(), "min eig", ev[0].item())
        s_keep = s.clone()
        linv = torch.full((l, l), float("nan"), device=dev)
        shift = (3e-4, 1e-6, 1e-7)[p]
        ops.cholesky_inverse([s], [linv], shift, 1e-12)
        torch.cuda.synchronize()
        eye = linv.double() @ s_keep.double() @ linv.double().t()
        print(f" pass {p}: Linv nans", nans(linv), "|Linv S Linv^T - I|max", (eye - torch.eye(l, device=dev)).abs().max().item(), "shift", shift,
              "max|Linv|", linv.abs().max().item())
        lref = torch.linalg.cholesky(s_keep.double() + shift * torch.eye(l, device=dev, dtype=torch.float64))
        print(f" pass {p}: Linv vs torch inverse rel", rel(linv, torch.linalg.inv(lref)))
        li = [torch.empty(l, l, device=dev, dtype=bf) for _ in range(3)]
        ops.split_bf16(linv, *li)
        q = torch.full((l, n), float("nan"), device=dev)
        ops.gemm_grouped([ops.make_problem(li, [lh, lm, ll], q, M=l, N=n, K=l, b_mn_major=True, terms=ops.TERMS_6)])
        torch.cuda.synchronize()
        print(f" pass {p}: Q nans", nans(q), "rel", rel(q, linv.double() @ ys),
              "orth", (q.double() @ q.double().t() - torch.eye(l, device=dev)).abs().max().item())
        cur = q


if __name__ == "__main__":
    main()
