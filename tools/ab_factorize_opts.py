"""Same-box A/B of FactorizeOptions variants on the bench step (config 2, graph replay, 10 steps, two rounds).
usage: python tools/ab_factorize_opts.py '[{}, {"second_pass_min_pivot": 0.01}]'   Design aid, not a bench line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xkv_b200 import compress, factorize

variants = json.loads(sys.argv[1]) if len(sys.argv) > 1 else [{}]
c = bench.CONFIGS[2]
keys, vals = bench.make_cache(c, torch.device("cuda"))
graphs = []
for v in variants:
    o = factorize.FactorizeOptions(**v)
    compress.compress_groups(keys, vals, c["rank_k"], c["rank_v"], opts=o, num_streams=8)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        res = compress.compress_groups(keys, vals, c["rank_k"], c["rank_v"], opts=o, num_streams=8)
    graphs.append((v, g, res))
for rnd in range(2):
    for v, g, _ in graphs:
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"round": rnd, "opts": v, "ms_per_step": round(e0.elapsed_time(e1) / 10, 3)}), flush=True)
