"""One pack of a 4-layer group at 64K tokens (K side), bracketed by cudaProfilerStart/Stop (for ncu), then the projection
A = X V of the same matrix.  Prints CUDA-event times."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xkv_b200 import ops

S, G, H, D, R = 65536, 4, 8, 128, 512
layers = [torch.randn(1, S, H, D, device="cuda").bfloat16().transpose(1, 2) for _ in range(G)]   # HF layout: (1, H, S, D) views
n = G * H * D
v = torch.randn(R, n, device="cuda").bfloat16()
a = torch.empty(S, R, device="cuda", dtype=torch.bfloat16)
x = ops.pack_group(layers)
ops.gemm_grouped([ops.make_problem([x[0]], [v], a, M=S, N=R, K=n)])
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
torch.cuda.profiler.start()
ev[0].record()
x = ops.pack_group(layers, out=x)
ev[1].record()
ops.gemm_grouped([ops.make_problem([x[0]], [v], a, M=S, N=R, K=n)])
ev[2].record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
nbytes = 2 * G * S * H * D * 2
print(json.dumps({"pack_ms": ev[0].elapsed_time(ev[1]), "pack_GBps": nbytes / ev[0].elapsed_time(ev[1]) / 1e6,
                  "project_ms": ev[1].elapsed_time(ev[2]), "project_TFLOPs": 2.0 * S * R * n / ev[1].elapsed_time(ev[2]) / 1e9}))
