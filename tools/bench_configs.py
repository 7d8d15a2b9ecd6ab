"""Throughput of the other BASELINE.json configurations on one B200 (design aid; bench.py's line stays config 2).

  config 3: single-layer SVD baseline, 32 layers -> 32 K + 32 V matrices 65536 x 1024, rank 128 / 192 (HBM-bound shape)
  config 4: Llama-3.1-70B-shaped KV, xKV-8 at 128K: matrices 131072 x 8192, rank 1024 / 1536 (a sample of groups)
  config 5: DeepSeek-V2-Lite MLA latents, 4-layer groups at 32K: 32768 x 2048 (and the last 3-layer group x 1536), rank 512

Each case: factorize_batch on resident token-major matrices, CUDA events, 1 warm-up + 3 timed calls; GB/s of bf16 KV."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xkv_b200 import factorize

CASES = [
    ("config3_K", 65536, 1024, 128, 32), ("config3_V", 65536, 1024, 192, 32),
    ("config5_latents", 32768, 2048, 512, 6), ("config5_last_group", 32768, 1536, 512, 1),
    ("config4_K", 131072, 8192, 1024, 2), ("config4_V", 131072, 8192, 1536, 2),
]


def main():
    only = sys.argv[1:] or None
    for name, m, n, r, count in CASES:
        if only and name not in only:
            continue
        g = torch.Generator(device="cuda").manual_seed(1)
        xs = []
        for b in range(count):
            # low-rank-plus-noise input generated in place (no QR at these sizes): decaying column scales
            x = torch.randn(m, n, device="cuda", generator=g)
            x *= (torch.arange(1, n + 1, device="cuda") ** -0.7)
            xs.append(x.to(torch.bfloat16))
            del x
        bmax = 16 if n <= 4096 else 2
        ws = torch.empty(factorize.workspace_bytes(min(count, bmax), m, n, r), dtype=torch.uint8, device="cuda")

        def run():
            out = []
            for lo in range(0, count, bmax):
                out += factorize.factorize_batch(xs[lo:lo + bmax], r, workspace=ws)
            return out

        run()
        torch.cuda.synchronize()
        times = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fs = run()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = sorted(times)[1]
        nbytes = count * m * n * 2
        ok = all(bool(torch.isfinite(f.A).all()) for f in fs)
        print(json.dumps({"case": name, "matrices": count, "m": m, "n": n, "rank": r, "ms": round(ms, 2),
                          "ms_per_matrix": round(ms / count, 3), "GBps_of_KV": round(nbytes / ms / 1e6, 1),
                          "hbm_two_pass_floor_ms": round(2 * nbytes / 6546.6e9 * 1e3, 2),
                          "gram_flops_floor_ms": round(count * m * n * n / 1393.7e12 * 1e3, 2), "finite": ok}), flush=True)
        del xs, ws, fs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
