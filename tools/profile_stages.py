"""Stage timings (CUDA events recorded by the driver) of the factorisation at config 2: K batch, V batch, and the K + V
mixed batch.   python tools/profile_stages.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xkv_b200 import factorize

c = bench.CONFIGS[2]
keys, vals = bench.make_cache(c, torch.device("cuda"))
rk = [[factorize.layer_rows(t) for t in g] for g in keys]
rv = [[factorize.layer_rows(t) for t in g] for g in vals]
opts = factorize.FactorizeOptions(profile=True)
for name, groups, ranks in (("K x8", rk, 512), ("V x8", rv, 768), ("K+V x16", rk + rv, [512] * 8 + [768] * 8)):
    for _ in range(2):
        fs = factorize.factorize_groups(groups, ranks, opts)
    t = fs[0].timings
    print(json.dumps({"batch": name, "total_ms": round(sum(t.values()), 2), **{k: round(v, 2) for k, v in t.items()}}))
