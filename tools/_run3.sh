cd $GRAFT_REPO_ROOT
timeout 120 tools/probe_decode_scores 0 0 512 4096 > gpurun_out/r02c_pair_first.txt 2>&1; echo "rc=$?" >> gpurun_out/r02c_pair_first.txt
cat gpurun_out/r02c_pair_first.txt
timeout 300 python -m pytest tests/test_decode_gpu.py -x -q -k "pair" 2>&1 | tail -15 > gpurun_out/r02c_pair_tests.txt
cat gpurun_out/r02c_pair_tests.txt
o=gpurun_out/r02c_pair_bisect.jsonl
: > $o
timeout 60 tools/probe_decode_scores 0 0 512 >> $o
timeout 60 tools/probe_decode_scores 0 1 512 >> $o
for st in 3 5 7; do timeout 60 tools/probe_decode_scores 0 0 512 65536 $st >> $o; done
timeout 60 tools/probe_decode_scores 0 0 256 >> $o
cat $o
