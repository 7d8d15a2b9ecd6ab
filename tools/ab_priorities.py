"""Same-box A/B of the stream priorities of the bench step's eight chains (4 K chains, then 4 V chains; config 2, graph
replay, 10 steps each, two rounds).  Design aid, not a bench line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xkv_b200 import compress, factorize

c = bench.CONFIGS[2]
dev = torch.device("cuda")
keys, vals = bench.make_cache(c, dev)
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -5)
print(json.dumps({"priority_range": [lo, hi]}), flush=True)
VH = [-1, -1, -1, -1, -3, -3, -3, -3]
RR = [-1, -2, -3]
schemes = {
    "8 streams (default)": dict(num_streams=8),
    "6 streams": dict(num_streams=6),
    "12 streams": dict(num_streams=12),
    "16 streams": dict(num_streams=16),
    "8 streams, V -3 K -2": dict(num_streams=8, priorities=[-2, -2, -2, -2, -3, -3, -3, -3]),
    "8 streams, V -2 K -1": dict(num_streams=8, priorities=[-1, -1, -1, -1, -2, -2, -2, -2]),
    "16 streams, V -3 K -2": dict(num_streams=16, priorities=[-2] * 8 + [-3] * 8),
}
graphs = {}
runs = {}
for name, pr in schemes.items():
    def run(pr=pr):
        pr = dict(pr)
        factorize._STAGGER_MARK = pr.pop("mark", 1)
        return compress.compress_groups(keys, vals, c["rank_k"], c["rank_v"], **pr)
    runs[name] = run
    run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        res = run()
    graphs[name] = (g, res)
for rnd in range(2):
    for name, (g, _) in graphs.items():
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"round": rnd, "scheme": name, "ms_per_step": round(e0.elapsed_time(e1) / 10, 3)}),
              flush=True)

