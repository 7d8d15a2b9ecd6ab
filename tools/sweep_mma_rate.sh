# tcgen05.mma issue-rate sweep (tools/probe_mma_rate.cu): shapes and forms the decode scores kernel can use
cd $GRAFT_REPO_ROOT
o=gpurun_out/r02_mma_rate_probe.jsonl
: > $o
for form in 0 1 2 3; do
  for n in 16 128 256; do
    timeout 60 tools/probe_mma_rate verify $form $n >> $o 2>> gpurun_out/r02_mma_rate_probe.err || echo "{\"verify_failed\": [$form, $n], \"rc\": $?}" >> $o
  done
done
for form in 0 1 2 3; do
  for n in 16 32 64 128 256; do
    timeout 60 tools/probe_mma_rate rate $form $n 2048 >> $o 2>> gpurun_out/r02_mma_rate_probe.err || echo "{\"rate_failed\": [$form, $n], \"rc\": $?}" >> $o
  done
done
cat $o
