#!/usr/bin/env python
"""bench.py — headline benchmark of the xKV hot path on B200 (contract: see the task statement).

Workload (BASELINE.json configs[1]): Llama-3.1-8B-shaped KV (32 layers, 8 KV heads x 128), 64K context,
batch 1, xKV-4 (8 groups of 4 layers, rank_k 512 / rank_v 768).  One *step* = prefill compression of the
whole cache: gather every group's K and V into token-major matrices and factorise them (16 matrices of
65536 x 4096).  `value` is GB/s of bf16 KV consumed with inputs resident in HBM; `e2e` is the same through
host buffers (pinned H2D of the KV and D2H of the factors inside the timed region).  The decode-side
number (fused reconstruct + attention, tok/s) is reported beside it as `decode`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched with torchrun; each rank owns its own 8 groups (weak scaling over layer groups, no
data-path collective); time is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "KV GB/s compressed (prefill, xKV-4, Llama-3.1-8B 64K)"
LAYERS, GROUP, HEADS, HEAD_DIM = 32, 4, 8, 128
RANK_K, RANK_V = 512, 768


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="xkv_b200", choices=["xkv_b200", "reference"])
    ap.add_argument("--tokens", type=int, default=65536)
    ap.add_argument("--cpu-sample-tokens", type=int, default=4096)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--streams", type=int, default=6, help="CUDA streams the K / V batches are spread over")
    ap.add_argument("--e2e-chunk", type=int, default=1, help="layer groups per pipelined chunk of the host-buffer path")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from the host instead of replaying a CUDA graph")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own arithmetic (oracle port of fake_svd etc.)
# ----------------------------------------------------------------------------------------------
def cpu_reference_step(sample_tokens: int, seed: int = 0):
    """One bounded sample of the workload on the host cores: one 4-layer group (K rank 512 + V rank 768)
    at `sample_tokens` tokens through the oracle's grouped merge. Returns (seconds, bytes_of_kv)."""
    import torch
    from oracle import xkv_oracle as O
    from xkv_b200 import synthetic

    keys = synthetic.make_group_kv(GROUP, HEADS, sample_tokens, HEAD_DIM, 1.0, seed)
    vals = synthetic.make_group_kv(GROUP, HEADS, sample_tokens, HEAD_DIM, 0.5, seed + 1)
    t0 = time.perf_counter()
    O.merge_group(keys, vals, RANK_K, RANK_V)
    dt = time.perf_counter() - t0
    nbytes = 2 * GROUP * HEADS * sample_tokens * HEAD_DIM * 2
    return dt, nbytes


def run_reference_arm(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    for _ in range(args.warmup):
        cpu_reference_step(args.cpu_sample_tokens)
    times = []
    nbytes = 0
    for i in range(args.steps):
        dt, nbytes = cpu_reference_step(args.cpu_sample_tokens, seed=i)
        times.append(dt)
    total = sum(times)
    value = nbytes * len(times) / total / 1e9
    sample = (f"one 4-layer group (8 KV heads x 128) at {args.cpu_sample_tokens} tokens per step: oracle port of "
              f"grouped_layer_merging (torch.linalg.svd fp32, K rank {RANK_K} + V rank {RANK_V})")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {
        "workload": f"Llama-3.1-8B-shaped KV, {LAYERS} layers x {HEADS} KV heads x {HEAD_DIM}, {args.tokens} tokens, "
                    f"batch 1, xKV-{GROUP} ({LAYERS // GROUP} groups), rank_k {RANK_K} / rank_v {RANK_V}",
        "tokens": args.tokens, "groups_per_gpu": LAYERS // GROUP, "parallelism": f"layer-groups x{args.gpus}",
        "l2": "inputs (8.6 GB per step) exceed L2", "cuda_streams": args.streams,
    }


# ----------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------------
def run_xkv_arm(args):
    import torch
    import torch.distributed as dist

    from xkv_b200 import compress, factorize, ops, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S = args.tokens
    ng = LAYERS // GROUP

    # ---- synthetic KV, resident in HBM (seeded per rank: every rank owns different groups) ----
    keys, vals = [], []
    for g in range(ng):
        keys.append(synthetic.make_group_kv(GROUP, HEADS, S, HEAD_DIM, 1.0, 1234 + 100 * rank + g, device=dev))
        vals.append(synthetic.make_group_kv(GROUP, HEADS, S, HEAD_DIM, 0.5, 5678 + 100 * rank + g, device=dev))
    kv_bytes = 2 * LAYERS * S * HEADS * HEAD_DIM * 2

    opts = factorize.FactorizeOptions()

    graphed = None
    if not args.no_graph:
        graphed = compress.GraphedCompressor(keys, vals, RANK_K, RANK_V, opts=opts, num_streams=args.streams)

    def step():
        if graphed is not None:
            return graphed.replay()
        return compress.compress_groups(keys, vals, RANK_K, RANK_V, opts=opts, num_streams=args.streams)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    launches_per_step = None
    if graphed is not None:
        # a replayed graph launches the captured kernels without passing through the library's counter
        c0 = ops.launch_count()
        compress.compress_groups(keys, vals, RANK_K, RANK_V, opts=opts)
        launches_per_step = ops.launch_count() - c0
        barrier()
        launches0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.profiler.start()      # no-op unless run under `ncu --profile-from-start off` (launch list of the timed steps)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = ops.launch_count() - launches0
    if launches_per_step is not None:
        launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    ms_step = ms_total / args.steps
    value = world * kv_bytes / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel: the symmetric Gram GEMM (tcgen05), timed with CUDA events ----
    popts = factorize.FactorizeOptions(profile=True)
    xk = compress.pack_groups(keys)
    fk = factorize.factorize_batch(xk, RANK_K, popts)
    stages_k = fk[0].timings
    del fk
    n = GROUP * HEADS * HEAD_DIM
    gram_ms = stages_k["gram_gemm"]                 # one launch, ng matrices
    alg_flops = ng * float(S) * n * n               # symmetric half of 2*m*n^2 per matrix
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    achieved = alg_flops / (gram_ms * 1e-3) / 1e12
    traffic = None
    try:
        per_matrix = json.load(open(os.path.join(ROOT, "profiles", "gram_traffic.json"))).get("dram_bytes_per_matrix")
        if per_matrix is not None and S == 65536:
            traffic = per_matrix * ng      # one launch covers the ng matrices of the batch
    except Exception:
        pass
    roofline = {
        "kernel": "gemm_kernel<1,1> (Gram X^T X, symmetric tiles, tcgen05 M128 N256 K16)",
        "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s",
        "traffic": traffic, "launch_ms": gram_ms, "algorithmic_flops_per_launch": alg_flops,
        "stages_ms_k_batch": stages_k,
        "pipeline_algorithmic_frac": (4.0 * S * n * (RANK_K + RANK_V) * ng / (ms_step * 1e-3) / 1e12) / peak_tf,
    }
    del xk

    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
        "gpu_launches": launches, "roofline": roofline,
        "launch_mode": "cuda-graph replay" if graphed is not None else "host enqueue",
    }

    # ---- decode: fused reconstruct + RoPE + attention over the factors, all 32 layers = one token ----
    if not args.no_decode:
        import math

        cos, sin = synthetic.llama3_rope(S, HEAD_DIM, device=dev)
        cos, sin = cos[0].contiguous(), sin[0].contiguous()
        hq = 32
        gen = torch.Generator(device=dev).manual_seed(7)
        q = torch.randn(LAYERS, hq, HEAD_DIM, device=dev, generator=gen).bfloat16()
        kt = torch.randn(LAYERS, HEADS, 1, HEAD_DIM, device=dev, generator=gen).bfloat16()
        vt = torch.randn(LAYERS, HEADS, 1, HEAD_DIM, device=dev, generator=gen).bfloat16()
        ws = torch.empty(ops.decode_workspace_bytes(hq, S, 1, RANK_V) + 4096, dtype=torch.uint8, device=dev)
        o = torch.empty(hq, HEAD_DIM, dtype=torch.bfloat16, device=dev)
        hd = HEADS * HEAD_DIM

        def one_token():
            for l in range(LAYERS):
                gf = out[l // GROUP]
                i = l % GROUP
                ops.decode_attention(q[l], gf.key.A, gf.key.V[i * hd:(i + 1) * hd], gf.value.A,
                                     gf.value.V[i * hd:(i + 1) * hd], HEADS, cos, sin, kt[l], vt[l],
                                     1.0 / math.sqrt(HEAD_DIM), out=o, workspace=ws)

        for _ in range(3):
            one_token()
        barrier()
        l0 = ops.launch_count()
        one_token()
        launches_per_token = ops.launch_count() - l0
        # one decode token = 32 layers x 5 kernels: replayed as a CUDA graph (static q / factors / workspace), like the
        # compress step, unless --no-graph
        token_graph = None
        if not args.no_graph:
            torch.cuda.synchronize()
            token_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(token_graph):
                one_token()
            token_graph.replay()
        barrier()
        ntok = 8
        e0.record()
        for _ in range(ntok):
            if token_graph is not None:
                token_graph.replay()
            else:
                one_token()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_tok = t.item() / ntok
        flops_k = 2.0 * S * RANK_K * HEADS * HEAD_DIM * LAYERS          # K^ reconstruction, the tensor-bound part
        bytes_a = float(S) * (RANK_K + RANK_V) * 2 * LAYERS             # token factors streamed once per layer
        peak_hbm = float(peaks.get("hbm_gbs", 6550.0))
        line["decode"] = {
            "metric": "decode tok/s reconstructed (attention over the factored cache, 32 layers, batch 1, 64K context)",
            "tok_s": world * 1e3 / ms_tok, "ms_per_token": ms_tok, "us_per_layer": 1e3 * ms_tok / LAYERS,
            "equiv_dense_kv_GBps": world * LAYERS * 2.0 * S * HEADS * HEAD_DIM * 2 / (ms_tok * 1e-3) / 1e9,
            "gpu_launches_per_token": launches_per_token,
            "launch_mode": "cuda-graph replay" if token_graph is not None else "host enqueue",
            "roofline": {"bound": "tensor", "achieved": flops_k / (ms_tok * 1e-3) / 1e12, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": flops_k / (ms_tok * 1e-3) / 1e12 / peak_tf,
                         "hbm_frac": bytes_a / (ms_tok * 1e-3) / 1e9 / peak_hbm,
                         "note": "whole decode step (scores + softmax + P*A_v + combine) against the K^ reconstruction flops"},
        }
        del ws
        # context, not a target: the library attention (torch SDPA, GQA) over an UNCOMPRESSED bf16 cache of one layer of
        # the same shape -- what the reference's decode costs once its dense K^ / V^ exist (llama.py:58-69).  It reads
        # 268 MB per layer against the factored cache's 168 MB per layer (and 5.7x less resident memory).
        if rank == 0:
            try:
                import torch.nn.functional as F

                kd = torch.randn(1, HEADS, S, HEAD_DIM, device=dev, dtype=torch.bfloat16)
                vd = torch.randn(1, HEADS, S, HEAD_DIM, device=dev, dtype=torch.bfloat16)
                qd = torch.randn(1, hq, 1, HEAD_DIM, device=dev, dtype=torch.bfloat16)
                for _ in range(3):
                    F.scaled_dot_product_attention(qd, kd, vd, enable_gqa=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(32):
                    F.scaled_dot_product_attention(qd, kd, vd, enable_gqa=True)
                e1.record()
                torch.cuda.synchronize()
                line["decode"]["dense_sdpa_us_per_layer"] = 1e3 * e0.elapsed_time(e1) / 32
                line["decode"]["dense_sdpa_note"] = ("torch SDPA over an uncompressed bf16 cache of the same shape "
                                                     "(library kernel, reported for context)")
                del kd, vd, qd
            except Exception as ex:   # an older torch without enable_gqa: the context number is optional
                line["decode"]["dense_sdpa_note"] = f"not measured: {type(ex).__name__}"

    # ---- end to end through the public API with HOST buffers ----
    if not args.no_e2e:
        # the synthetic cache is moved to pinned host memory and the device copies are dropped: every e2e step
        # starts from HOST buffers and ends with the factors back in HOST buffers
        h_keys = [[t.transpose(1, 2).contiguous().cpu().pin_memory() for t in grp] for grp in keys]
        h_vals = [[t.transpose(1, 2).contiguous().cpu().pin_memory() for t in grp] for grp in vals]
        r_k, r_v, n_cols = RANK_K, RANK_V, GROUP * HEADS * HEAD_DIM
        h_out = []
        for _ in range(ng):
            for r in (r_k, r_v):
                h_out.append(torch.empty(S, r, dtype=torch.bfloat16).pin_memory())
                h_out.append(torch.empty(r, n_cols, dtype=torch.bfloat16).pin_memory())
        d2h_bytes = sum(t.numel() * t.element_size() for t in h_out)
        if graphed is not None:
            del graphed, out
            graphed = None
        del keys, vals
        torch.cuda.empty_cache()

        def e2e_step():
            compress.compress_groups_from_host(h_keys, h_vals, RANK_K, RANK_V, dev, chunk_groups=args.e2e_chunk,
                                               opts=opts, host_out=h_out)
            return d2h_bytes

        d2h = e2e_step()
        barrier()
        n_e2e = min(args.steps, 3)
        e0.record()
        for _ in range(n_e2e):
            d2h = e2e_step()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item() / n_e2e
        line["e2e"] = {"value": world * kv_bytes / (ms_e2e * 1e-3) / 1e9, "unit": "GB/s",
                       "h2d_bytes_per_step": kv_bytes, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
                       "steps": n_e2e}
        del h_keys, h_vals, h_out

    if rank == 0:
        line["clocks"] = clocks
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            dt, nbytes = cpu_reference_step(args.cpu_sample_tokens)
            line["cpu_baseline"] = {
                "value": nbytes / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                "sample": f"one 4-layer group at {args.cpu_sample_tokens} tokens (K rank {RANK_K} + V rank {RANK_V}) "
                          f"through the oracle port of grouped_layer_merging, {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_xkv_arm(args)


if __name__ == "__main__":
    main()
