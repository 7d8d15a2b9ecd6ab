cd $GRAFT_REPO_ROOT
o=gpurun_out/r02_tma_ring_probe.jsonl
: > $o
for st in 3 5 8 12; do timeout 60 tools/probe_tma_ring $st 128 8 1 0 >> $o; done
for st in 3 6; do timeout 60 tools/probe_tma_ring $st 256 8 1 0 >> $o; done
for cl in 2 4 8; do timeout 60 tools/probe_tma_ring 5 128 8 $cl 0 >> $o; timeout 60 tools/probe_tma_ring 12 128 8 $cl 0 >> $o; done
timeout 60 tools/probe_tma_ring 5 128 1 1 0 524288 >> $o
timeout 60 tools/probe_tma_ring 12 128 1 1 0 524288 >> $o
for h in 256 512; do timeout 60 tools/probe_tma_ring 5 128 8 1 $h >> $o; timeout 60 tools/probe_tma_ring 12 128 8 1 $h >> $o; timeout 60 tools/probe_tma_ring 12 128 8 2 $h >> $o;  done
