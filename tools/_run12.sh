cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_compress_gpu.py tests/test_cache_gpu.py tests/test_factorize_gpu.py -x -q 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e'], d['clocks'])"; done
