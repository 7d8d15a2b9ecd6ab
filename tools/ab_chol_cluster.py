"""Same-box A/B of the Cholesky cluster size per matrix (8 automatic, capped at 4 / 2) on the bench step (config 2, graph
replay, 10 steps, two rounds) and on a lone layer group.  Design aid, not a bench line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xkv_b200 import _lib, compress

lib = _lib.load()
c = bench.CONFIGS[2]
keys, vals = bench.make_cache(c, torch.device("cuda"))
graphs = []
for cap in (0, 4, 2):
    for name, k, v in (("8 groups", keys, vals), ("1 group", keys[:1], vals[:1])):
        lib.xkv_cholesky_set_cluster_cap(cap)
        compress.compress_groups(k, v, c["rank_k"], c["rank_v"], num_streams=8)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            res = compress.compress_groups(k, v, c["rank_k"], c["rank_v"], num_streams=8)
        graphs.append((cap, name, g, res))
lib.xkv_cholesky_set_cluster_cap(0)
for rnd in range(2):
    for cap, name, g, _ in graphs:
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"round": rnd, "cluster_cap": cap, "work": name, "ms_per_step": round(e0.elapsed_time(e1) / 10, 3)}), flush=True)
