cd $GRAFT_REPO_ROOT
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-decode --no-extras"
for rep in 1 2; do
for cfg in "5 6" "6 6" "4 6" "7 6" "8 6" "9 6" "4 4" "4 8" "7 8" "4 12"; do
  set -- $cfg
  XKV_STREAM_PRIO=$1 timeout 300 $B --streams $2 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('prio_mode', $1, 'streams', $2, round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
done
