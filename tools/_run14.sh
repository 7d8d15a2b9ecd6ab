cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_cache_gpu.py tests/test_generate_gpu.py -x -q 2>&1 | tail -4
for v in 0 4 0 4; do timeout 120 python tools/run_decode_once.py 65536 8 --graph --variant $v 2>/dev/null | tail -1; done
