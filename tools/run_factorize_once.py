"""One warm-up call, then one factorize_batch call bracketed by cudaProfilerStart/Stop (for ncu
--profile-from-start off). Prints the CUDA-event time of the profiled call."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xkv_b200 import factorize, ops, synthetic

S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 512
n = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
xs = [synthetic.group_matrix(S, n, 1.0, seed=b, device="cuda") for b in range(B)]
ws = torch.empty(factorize.workspace_bytes(min(B, 16), S, n, rank), dtype=torch.uint8, device="cuda")
factorize.factorize_batch(xs, rank, workspace=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = ops.launch_count()
torch.cuda.profiler.start()
e0.record()
factorize.factorize_batch(xs, rank, workspace=ws)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(json.dumps({"S": S, "B": B, "rank": rank, "n": n, "ms": e0.elapsed_time(e1), "launches": ops.launch_count() - l0}))
