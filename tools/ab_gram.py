"""A/B of the Gram stage (CUDA events of the driver's stage marks) for accumulation-piece lengths. Design aid."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xkv_b200 import factorize, synthetic
xs = [synthetic.group_matrix(65536, 4096, 1.0, seed=b, device="cuda") for b in range(8)]
for rep in range(3):
    for chunk in (0, 8192, 16384, 32768):
        o = factorize.FactorizeOptions(profile=True, gram_chunk_tokens=chunk)
        fs = factorize.factorize_batch(xs, 512, o)
        t = fs[0].timings
        print(json.dumps({"chunk": chunk, "gram_ms": round(t["gram_gemm"], 3), "reduce_split_ms": round(t["gram_reduce_split"], 3),
                          "sum_ms": round(sum(t.values()), 2)}), flush=True)
