"""One layer group (K + V) on one GPU: separate chains on two streams vs one mixed-rank driver call, CUDA-graph replay.
    python tools/probe_lone_group.py [groups]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xkv_b200 import compress

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 1
c = bench.CONFIGS[2]
keys, vals = bench.make_cache(c, torch.device("cuda"))
keys, vals = keys[:ng], vals[:ng]
res = {"groups": ng}
for name, kw in (("separate_2streams", dict(mixed=False, num_streams=2)), ("separate_1stream", dict(mixed=False, num_streams=1)),
                 ("mixed", dict(mixed=True, num_streams=1))):
    for _ in range(2):
        compress.compress_groups(keys, vals, c["rank_k"], c["rank_v"], **kw)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        compress.compress_groups(keys, vals, c["rank_k"], c["rank_v"], **kw)
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    res[name + "_ms"] = round(e0.elapsed_time(e1) / 10, 3)
print(json.dumps(res))
