# Same-box A/B of two builds of the library on the bench step (config 2, graph replay): tools/ab/libxkv_prev.so against the
# in-tree build, alternating, one process each.   usage: bash tools/ab_lib_step.sh [rounds]
# (put the build to compare against there first: cp xkv_b200/libxkv_b200.so tools/ab/libxkv_prev.so before the change;
# *.so files are git-ignored and travel to the GPU box with the snapshot)
cd $GRAFT_REPO_ROOT
for i in $(seq 1 ${1:-2}); do
  XKV_B200_LIB=$PWD/tools/ab/libxkv_prev.so python tools/ab_factorize_opts.py '[{}]' | sed 's/^/prev /'
  python tools/ab_factorize_opts.py '[{}]' | sed 's/^/new  /'
done
