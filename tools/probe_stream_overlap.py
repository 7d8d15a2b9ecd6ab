"""Do the K and V factorisation chains of one compress step really overlap on the device?  CUDA events at the start and
end of every job (a batch of matrices on its own stream), printed relative to the step's start.
    python tools/probe_stream_overlap.py [streams]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xkv_b200 import compress

streams = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = bench.CONFIGS[2]
keys, vals = bench.make_cache(c, torch.device("cuda"))
for _ in range(2):
    compress.compress_groups(keys, vals, c["rank_k"], c["rank_v"], num_streams=streams)
torch.cuda.synchronize()
ev = []
t0 = torch.cuda.Event(enable_timing=True)
t0.record()
compress.compress_groups(keys, vals, c["rank_k"], c["rank_v"], num_streams=streams, job_events=ev)
t1 = torch.cuda.Event(enable_timing=True)
t1.record()
torch.cuda.synchronize()
print(json.dumps({"streams": streams, "step_ms": round(t0.elapsed_time(t1), 2),
                  "jobs": [{"job": j, "rank": r, "matrices": n, "start_ms": round(t0.elapsed_time(a), 2),
                            "end_ms": round(t0.elapsed_time(b), 2)} for j, r, n, a, b in ev]}))
