cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_decode_gpu.py -x -q 2>&1 | tail -3 > gpurun_out/r02b_decode_tests.txt
cat gpurun_out/r02b_decode_tests.txt
o=gpurun_out/r02b_decode_scores_bisect.jsonl
: > $o
for d in 0 4 68 64 580; do timeout 60 tools/probe_decode_scores $d 1 512 >> $o; done
timeout 60 tools/probe_decode_scores 0 1 256 >> $o
cat $o
timeout 120 python tools/run_decode_once.py 65536 8 --graph 2>&1 | tail -2
