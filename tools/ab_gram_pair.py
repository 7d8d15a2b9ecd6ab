"""Same-box A/B of the Gram launch of the bench step (8 matrices 65536 x 4096, four layer tensors each, read in place,
accumulation pieces of 16384 tokens): CTA pairs (gram_pair_kernel) against the single-CTA 128 x 256 tiles.  CUDA events on
the launching stream, L2 flushed by the operands themselves (4.3 GB per launch).  Design aid, not a bench line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xkv_b200 import _lib, ops

S, n, B, G = 65536, 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 8, 4
lib = _lib.load()
layers = [[torch.randn(S, n // G, device="cuda").bfloat16() for _ in range(G)] for _ in range(B)]
slabs = torch.empty(B, n, n, device="cuda", dtype=torch.float32)
probs = [ops.make_problem([], [], slabs[b], M=n, N=n, K=S, a_mn_major=True, b_mn_major=True, sym_upper=True,
                          accum_phases=4, a_layers=layers[b], b_layers=layers[b]) for b in range(B)]
flop = B * S * n * n
res = {}
for rep in range(3):
    for pair in (6, 7, 5, 4, 0):
        lib.xkv_gemm_set_gram_pair(pair)
        for _ in range(2):
            ops.gemm_grouped(probs)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm_grouped(probs)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        if rep == 0:
            res[pair] = slabs.clone()
        print(json.dumps({"pair": pair, "matrices": B, "ms_min": round(min(ts), 4), "ms_avg": round(sum(ts) / len(ts), 4),
                          "TFLOPs_useful": round(flop / min(ts) / 1e9, 1)}), flush=True)
lib.xkv_gemm_set_gram_pair(1)
mask = torch.triu(torch.ones(n, n, dtype=torch.bool, device="cuda"))
print(json.dumps({"bit_identical_upper": bool(all(torch.equal(res[p][b][mask], res[0][b][mask]) for b in range(B) for p in res))}))
