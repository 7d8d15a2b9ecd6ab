# Final evidence of round 2 (each ncu run only after the same command exited 0 without it): driver-style bench line, launch
# list of one timed step, launch list of four decode layers.
set -x
cd $GRAFT_REPO_ROOT
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-decode --no-extras"
$B > gpurun_out/r02_final_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2500 --csv --log-file gpurun_out/r02_bench_step_launches_last.csv $B > /dev/null 2>&1
D="python tools/run_decode_once.py 65536 4"
$D > gpurun_out/r02_final_plain_decode.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 100 --csv --log-file gpurun_out/r02_decode4_launches_last.csv $D > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_bench_step_launches_last.csv | head -14
python tools/summarize_launches.py gpurun_out/r02_decode4_launches_last.csv | head -12
head -c 400 gpurun_out/r02_final_bench_n1.json
