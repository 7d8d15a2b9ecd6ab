#!/bin/bash
# bisect of the transposed decode scores kernel: XKV_DECODE_DBG bit 0 no reconstruction MMAs, 1 no K^rot stores,
# 2 no score MMAs, 3 no RoPE (no cos/sin loads, no shuffles), 4 no TMEM loads in the epilogue
for d in ${@:-0 1 2 4 8 16 10 26 27 31}; do
  XKV_VARIANT=4 XKV_DECODE_DBG=$d ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:decode_scores --csv python tools/run_decode_once.py 65536 2 2>/dev/null | grep decode_scores | awk -F'","' -v d=$d '{gsub(/"/,"",$NF); s+=$NF; n++} END {print "dbg", d, s/n/1000, "us"}'
done
