cd $GRAFT_REPO_ROOT
C="python tools/run_compress_once.py"
$C && ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kernel -c 1 -o gpurun_out/r02_gram_inplace_full $C > /dev/null 2>&1
ls -la gpurun_out/r02_gram_inplace_full.ncu-rep
