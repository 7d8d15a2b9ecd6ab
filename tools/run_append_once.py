"""A group's K and V append projection (config 2 shapes), for ncu: python tools/run_append_once.py [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xkv_b200 import ops

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = 4096
xs = [torch.randn(T, n, device="cuda").bfloat16() for _ in range(2)]
vs = [torch.randn(n, r, device="cuda").bfloat16() for r in (512, 768)]
outs = [torch.empty(T, r, dtype=torch.bfloat16, device="cuda") for r in (512, 768)]
ws = torch.empty(1 << 24, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    ops.append_project_many(xs, vs, outs, workspace=ws)
torch.cuda.synchronize()
print("ok")
