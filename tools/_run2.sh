cd $GRAFT_REPO_ROOT
o=gpurun_out/r02b_decode_scores_bisect2.jsonl
: > $o
for d in 0 16 8 24 4 20; do timeout 60 tools/probe_decode_scores $d 1 512 >> $o; done
for st in 2 3 4; do timeout 60 tools/probe_decode_scores 0 1 512 65536 $st >> $o; done
for S in 16384 32768 131072; do timeout 60 tools/probe_decode_scores 0 1 512 $S >> $o; done
cat $o
