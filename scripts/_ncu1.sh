cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cap() {  # name regex skip count
  timeout 280 python scripts/ncu_targets.py $1 > gpurun_out/plain_$1.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/r2_$1 python scripts/ncu_targets.py $1 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"; tail -2 gpurun_out/ncu_$1.log
}
cap step bprmf_step_kernel 3 2
export WR_EPOCH_KERNEL=two_barrier      # the first resident kernel (scripts/_ncu2.sh captures the owner-computes one)
cap epoch bprmf_epoch_kernel 1 1
unset WR_EPOCH_KERNEL
cap adam_big 'adam_sweep_kernel|bpr_fwd_bwd_kernel' 2 2
cap eval_tc64 eval_tc_rank_kernel 1 1
cap eval_tc128 eval_tc_rank_kernel 1 1
cap spmm_cfg4 csr_spmm_kernel 1 1
ls -la gpurun_out/*.ncu-rep
