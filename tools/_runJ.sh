cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_decode_gpu.py -x -q 2>&1 | tail -3
for i in 1 2; do timeout 60 tools/probe_decode_scores 0 0 512; done
timeout 60 tools/probe_decode_scores 0 0 512 65536 9
timeout 60 tools/probe_decode_scores 0 0 256
for i in 1 2; do timeout 120 python tools/run_decode_once.py 65536 8 --graph | tail -1; done
