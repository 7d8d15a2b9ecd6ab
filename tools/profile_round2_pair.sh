# Launch list of the bench step and one full ncu capture of the CTA-pair Gram (each ncu run only after the same command
# exited 0 without it)
set -x
cd $GRAFT_REPO_ROOT
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-decode --no-extras"
$B > gpurun_out/r02s_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2500 --csv --log-file gpurun_out/r02_bench_step_launches_pair.csv $B > /dev/null 2>&1
C="python tools/run_compress_once.py"
$C && ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_pair_kernel -c 1 -o gpurun_out/r02_gram_pair_full $C > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_bench_step_launches_pair.csv | head -30
ls -la gpurun_out/r02_gram_pair_full.ncu-rep
