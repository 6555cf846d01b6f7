cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_full_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_full_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_b2.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_ref.json 2>/dev/null; echo "ref rc=$?"
