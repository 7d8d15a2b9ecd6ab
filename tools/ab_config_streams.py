"""Graph-replay time of a configuration's compress step against the number of chains (CUDA streams) its K / V matrices are
spread over, plus one host-enqueued step with per-chain start / end events.  usage: ab_config_streams.py CONFIG [streams ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xkv_b200 import compress

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 3
streams = [int(x) for x in sys.argv[2:]] or [8, 16, 32]
c = bench.CONFIGS[cid]
dev = torch.device("cuda")
keys, vals = bench.make_cache(c, dev)
sizes = bench.group_sizes(c)
full = [g for g, s in enumerate(sizes) if s == sizes[0]]
keys, vals = [keys[g] for g in full], [vals[g] for g in full]


def run(ns, ev=None):
    return compress.compress_groups(keys, vals, c["rank_k"], c["rank_v"], merge_value=c["merge_value"], num_streams=ns, job_events=ev)


graphs = {}
for ns in streams:
    run(ns)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        res = run(ns)
    graphs[ns] = (g, res)
for rnd in range(2):
    for ns, (g, _) in graphs.items():
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"config": cid, "groups": len(keys), "round": rnd, "streams": ns, "ms_per_step": round(e0.elapsed_time(e1) / 10, 3)}), flush=True)
ev = []
t0 = torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
t0.record()
run(streams[0], ev)
torch.cuda.synchronize()
print(json.dumps({"config": cid, "streams": streams[0], "mode": "host-enqueued",
                  "jobs": [[j, r, n, round(t0.elapsed_time(a), 2), round(t0.elapsed_time(b), 2)] for j, r, n, a, b in ev]}))
