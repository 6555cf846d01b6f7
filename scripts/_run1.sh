set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "resident or host_fed or epoch_call or bprmf" > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
tail -15 gpurun_out/r2_t1.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err; echo "bench rc=$?"
cat gpurun_out/r2_b1.json | head -c 3000; tail -5 gpurun_out/r2_b1.err
