"""One decode token (32 layers) over random factors, bracketed by cudaProfilerStart/Stop (for ncu)."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xkv_b200 import ops, synthetic

S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
H, D, HQ, RK, RV, G = 8, 128, 32, 512, 768, 4
n = G * H * D
dev = "cuda"
a_k = torch.randn(S, RK, device=dev).bfloat16()
a_v = torch.randn(S, RV, device=dev).bfloat16()
v_k = torch.randn(n, RK, device=dev).bfloat16()
v_v = torch.randn(n, RV, device=dev).bfloat16()
cos, sin = synthetic.llama3_rope(S, D, device=dev)
cos, sin = cos[0].contiguous(), sin[0].contiguous()
rope_t = None if os.environ.get("XKV_NO_ROPE_T") else ops.rope_tables_dim_major(cos, sin)
q = torch.randn(HQ, D, device=dev).bfloat16()
kt = torch.randn(H, 1, D, device=dev).bfloat16()
vt = torch.randn(H, 1, D, device=dev).bfloat16()
if os.environ.get("XKV_VARIANT"):
    from xkv_b200 import _lib
    _lib.load().xkv_decode_set_variant(int(os.environ["XKV_VARIANT"]))
ws = torch.empty(ops.decode_workspace_bytes(HQ, S, 1, RV) + 4096, dtype=torch.uint8, device=dev)


def run():
    for l in range(layers):
        i = l % G
        ops.decode_attention(q, a_k, v_k[i * H * D:(i + 1) * H * D], a_v, v_v[i * H * D:(i + 1) * H * D], H, cos, sin,
                             kt, vt, 1.0 / math.sqrt(D), workspace=ws, rope_t=rope_t)


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
run()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(json.dumps({"S": S, "layers": layers, "ms": e0.elapsed_time(e1), "us_per_layer": 1e3 * e0.elapsed_time(e1) / layers}))
