set -x
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-decode --no-extras"
$B > gpurun_out/r02_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file gpurun_out/r02_bench_step_launches.csv $B > /dev/null 2>&1
D="python tools/run_decode_once.py 65536 4"
$D > gpurun_out/r02_plain_decode.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 100 --csv --log-file gpurun_out/r02_decode4_launches.csv $D > /dev/null 2>&1
$D > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_scores -c 1 -o gpurun_out/r02_scores_full $D > /dev/null 2>&1
A="python tools/run_append_once.py 1"
$A > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:append -s 2 -c 1 -o gpurun_out/r02_append_full $A > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_*launches.csv
