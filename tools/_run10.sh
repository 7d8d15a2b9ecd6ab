cd $GRAFT_REPO_ROOT
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-decode --no-extras"
for rep in 1 2; do
for m in 0 1 2 3 4; do
  XKV_STREAM_PRIO=$m timeout 300 $B 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('prio_mode', $m, d['ms_per_step'], d['clocks']['sm_mhz'])"
done
done
XKV_STREAM_PRIO=1 timeout 300 $B --no-graph 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('prio_mode 1 nograph', d['ms_per_step'])"
XKV_STREAM_PRIO=0 timeout 300 $B --no-graph 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('prio_mode 0 nograph', d['ms_per_step'])"
