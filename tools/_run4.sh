cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_cache_gpu.py tests/test_generate_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/r02c_tests.txt
cat gpurun_out/r02c_tests.txt
o=gpurun_out/r02c_pair_bisect2.jsonl
: > $o
for st in 6 7 8 9; do timeout 60 tools/probe_decode_scores 0 0 512 65536 $st >> $o; done
for S in 16384 32768 131072; do timeout 60 tools/probe_decode_scores 0 0 512 $S >> $o; done
timeout 60 tools/probe_decode_scores 0 0 768 65536 >> $o
timeout 60 tools/probe_decode_scores 0 0 1024 65536 >> $o
cat $o
timeout 120 python tools/run_decode_once.py 65536 8 --graph 2>&1 | tail -2
