"""When does each of the bench step's eight chains start and end inside a live step?  (config 2, launches enqueued from the
host, CUDA events around every job; times relative to the step's first event.)  Design aid."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xkv_b200 import compress

c = bench.CONFIGS[2]
dev = torch.device("cuda")
keys, vals = bench.make_cache(c, dev)
for rep in range(4):
    ev = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    compress.compress_groups(keys, vals, c["rank_k"], c["rank_v"], num_streams=8, job_events=ev)
    t1.record()
    torch.cuda.synchronize()
    if rep >= 2:
        print(json.dumps({"step_ms": round(t0.elapsed_time(t1), 2),
                          "jobs": [{"job": j, "rank": r, "groups": n, "start_ms": round(t0.elapsed_time(a), 2),
                                    "end_ms": round(t0.elapsed_time(b), 2)} for j, r, n, a, b in ev]}), flush=True)
