"""Ad-hoc kernel timing on a B200 (CUDA events on the launching stream). Not a bench line."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xkv_b200 import ops


def timeit(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    G, H, D = 4, 8, 128
    n = G * H * D
    layers = [torch.randn(1, S, H, D, device="cuda").bfloat16().transpose(1, 2) for _ in range(G)]
    X = torch.empty(1, S, n, device="cuda", dtype=torch.bfloat16)
    best, avg = timeit(lambda: ops.pack_group(layers, out=X))
    byt = 2 * S * n * 2
    print(json.dumps({"kernel": "pack", "S": S, "n": n, "ms_best": best, "ms_avg": avg, "GBps": byt / best / 1e6}))
    X2 = X[0]
    for split in (1, 2, 4, 8):
        slabs = torch.empty(split, n, n, device="cuda", dtype=torch.float32)
        p = ops.make_problem([X2], [X2], slabs[0], M=n, N=n, K=S, a_mn_major=True, b_mn_major=True,
                             sym_upper=True, split_k=split, split_stride=slabs.stride(0))
        best, avg = timeit(lambda: ops.gemm_grouped([p]))
        fl = S * n * n  # symmetric half of 2*S*n*n
        print(json.dumps({"kernel": "gram_sym", "split": split, "ms_best": best, "ms_avg": avg,
                          "useful_TFLOPs": fl / best / 1e9, "executed_TFLOPs": 2.0 * S * 128 * 256 * 272 / best / 1e9}))
        del slabs
    for r in (512, 768):
        V = torch.randn(r, n, device="cuda").bfloat16()
        A = torch.empty(S, r, device="cuda", dtype=torch.bfloat16)
        p = ops.make_problem([X2], [V], A, M=S, N=r, K=n)
        best, avg = timeit(lambda: ops.gemm_grouped([p]))
        print(json.dumps({"kernel": "project", "r": r, "ms_best": best, "ms_avg": avg,
                          "TFLOPs": 2.0 * S * n * r / best / 1e9}))
    # power step shape: Yt[l x n] = Qt[l x n] * G (K-major x K-major), 3 terms
    l = 576
    Qh = [torch.randn(l, n, device="cuda").bfloat16() for _ in range(2)]
    Gh = [torch.randn(n, n, device="cuda").bfloat16() for _ in range(2)]
    Y = torch.empty(l, n, device="cuda")
    p = ops.make_problem(Qh, Gh, Y, M=l, N=n, K=n, terms=ops.TERMS_3)
    best, avg = timeit(lambda: ops.gemm_grouped([p]))
    print(json.dumps({"kernel": "power_step_T3", "l": l, "ms_best": best, "TFLOPs": 3 * 2.0 * l * n * n / best / 1e9}))
    # reference point: torch matmul (cuBLAS) on the projection shape
    V = torch.randn(n, 512, device="cuda").bfloat16()
    best, avg = timeit(lambda: torch.matmul(X2, V))
    print(json.dumps({"kernel": "cublas_project_r512", "ms_best": best, "TFLOPs": 2.0 * S * n * 512 / best / 1e9}))
    Xf = X2[:, :]
    best, avg = timeit(lambda: torch.matmul(Xf.t(), Xf))
    print(json.dumps({"kernel": "cublas_gram_full", "ms_best": best, "TFLOPs": 2.0 * S * n * n / best / 1e9}))


if __name__ == "__main__":
    main()
