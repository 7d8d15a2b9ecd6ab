# second sweep: what makes the N = 128 MMAs of the decode scores kernel slower than the 64 clk measured with a tiny B footprint?
cd $GRAFT_REPO_ROOT
o=gpurun_out/r02_mma_rate_probe2.jsonl
: > $o
run() { timeout 60 tools/probe_mma_rate rate "$@" >> $o 2>> gpurun_out/r02_mma_rate_probe.err || echo "{\"failed\": \"$*\", \"rc\": $?}" >> $o; }
# form N mmas b_blocks a_stages commit_every random
run 0 128 2048 1 4 0 0
run 0 128 2048 8 4 0 0
run 0 128 2048 8 5 0 0
run 0 128 2048 8 5 1 0
run 0 128 2048 8 5 1 1
run 0 128 2048 1 4 0 1
run 0 128 2048 8 5 0 1
run 0 256 2048 1 4 0 1
run 0 256 2048 4 5 0 1
run 2 128 2048 8 5 1 1
run 2 128 2048 16 5 1 1
run 2 256 2048 8 5 1 1
run 1 128 2048 8 5 1 1
run 0 128 8192 8 5 1 1
cat $o
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
