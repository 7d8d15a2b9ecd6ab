cd $GRAFT_REPO_ROOT
o=gpurun_out/r02c_decode_layer_stages.jsonl
: > $o
for st in 0 5 6 7 8; do timeout 120 python tools/run_decode_once.py 65536 8 --graph --stages $st 2>/dev/null | tail -1 >> $o; done
timeout 120 python tools/run_decode_once.py 65536 8 --graph --variant 2 2>/dev/null | tail -1 >> $o
cat $o
D="python tools/run_decode_once.py 65536 4"
$D > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 100 --csv --log-file gpurun_out/r02c_decode4_launches.csv $D > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02c_decode4_launches.csv
