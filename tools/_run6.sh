cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_factorize_gpu.py tests/test_compress_gpu.py -x -q 2>&1 | tail -4 > gpurun_out/r02d_tests.txt
cat gpurun_out/r02d_tests.txt
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-decode --no-extras"
for rep in 1 2; do
  XKV_B200_LIB=$PWD/tools/ab/libxkv_old_gemm.so timeout 300 $B 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('old', d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['stages_ms_k_batch'])"
  timeout 300 $B 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('new', d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['stages_ms_k_batch'])"
done
