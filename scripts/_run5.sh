cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "spmm or lightgcn or propagation" 2>&1 | tail -4
timeout 280 python scripts/ncu_targets.py spmm_cfg4 > gpurun_out/plain_spmm2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_spmm_kernel -s 1 -c 1 -f -o gpurun_out/r2_spmm_cfg4_hot python scripts/ncu_targets.py spmm_cfg4 > gpurun_out/ncu_spmm2.log 2>&1
echo "ncu rc=$?"
