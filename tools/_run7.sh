cd $GRAFT_REPO_ROOT
timeout 900 python tools/sweep_fullsize.py '[{"power_iters": 4}, {"power_iters": 3}, {"power_iters": 2}, {"power_iters": 1}]' 2>&1 | tail -6 | tee gpurun_out/r02d_power_iters_sweep.jsonl
