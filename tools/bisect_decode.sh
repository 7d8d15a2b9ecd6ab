#!/bin/bash
# On-device bisect of the decode scores kernel (tools/probe_decode_scores.cu; bits: 1 no reconstruction MMAs, 2 stages
# released by a plain arrive, 4 no epilogue pipeline, 8 no RoPE, 16 epilogue stops after reading the accumulator)
cd "$(dirname "$0")/.."
out=${OUT:-gpurun_out/r02_decode_scores_bisect.jsonl}
: > $out
for d in ${@:-0 8 16 4 5 7 3}; do timeout 60 tools/probe_decode_scores $d 1 512 >> $out; done
for d in 0 4; do timeout 60 tools/probe_decode_scores $d 1 256 >> $out; done
cat $out
