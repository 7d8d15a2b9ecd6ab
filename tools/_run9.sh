cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_compress_gpu.py -x -q 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-decode --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'])"; done
