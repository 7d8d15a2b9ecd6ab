# Launch lists and one full ncu capture of the final round-2 code (each ncu run only after the same command exited 0 without it)
set -x
cd $GRAFT_REPO_ROOT
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-decode --no-extras"
$B > gpurun_out/r02i_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2500 --csv --log-file gpurun_out/r02_bench_step_launches_final.csv $B > /dev/null 2>&1
D="python tools/run_decode_once.py 65536 4"
$D > gpurun_out/r02i_plain_decode.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 100 --csv --log-file gpurun_out/r02_decode4_launches_final.csv $D > /dev/null 2>&1
$D > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_scores -c 1 -o gpurun_out/r02_scores_pair_final $D > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_bench_step_launches_final.csv | head -30
python tools/summarize_launches.py gpurun_out/r02_decode4_launches_final.csv
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-600
