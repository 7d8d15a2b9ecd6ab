"""One warm-up call, then one compress_groups call on ONE layer group read in place (per-layer tensor maps), bracketed by
cudaProfilerStart/Stop (for ncu --profile-from-start off).  Prints the CUDA-event time of the profiled call."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xkv_b200 import compress, ops

c = bench.CONFIGS[2]
keys, vals = bench.make_cache(c, torch.device("cuda"))
keys, vals = keys[:1], vals[:1]
compress.compress_groups(keys, vals, c["rank_k"], c["rank_v"], num_streams=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = ops.launch_count()
torch.cuda.profiler.start()
e0.record()
compress.compress_groups(keys, vals, c["rank_k"], c["rank_v"], num_streams=1)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(json.dumps({"groups": 1, "ms": e0.elapsed_time(e1), "launches": ops.launch_count() - l0}))
