cd $GRAFT_REPO_ROOT
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b3.json 2> gpurun_out/r2_b3.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_b3.err
