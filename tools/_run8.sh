cd $GRAFT_REPO_ROOT
timeout 1200 python bench.py --gpus 1 --steps 10 --warmup 5 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02e_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['decode']['tok_s'], d['decode']['us_per_layer'], d['clocks'])
print({k:(v.get('value'), v.get('decode',{}).get('tok_s')) for k,v in d['other_configs'].items()})
PY
D="python tools/run_decode_once.py 65536 4"
$D > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_scores -c 1 -o gpurun_out/r02e_scores_pair_full $D > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
