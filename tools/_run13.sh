cd $GRAFT_REPO_ROOT
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-decode --no-extras"
for rep in 1 2; do
for o in '{}' '{"gram_split_k": 2}' '{"gram_split_k": 4}' '{"want_sigma": false}' '{"small_split_k": 4}' '{"small_split_k": 16}'; do
  timeout 300 $B --factorize-opts "$o" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$o', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
done
