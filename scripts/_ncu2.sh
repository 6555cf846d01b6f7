cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cap() {  # name regex skip count
  timeout 280 python scripts/ncu_targets.py $1 > gpurun_out/plain_$1.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/r2b_$1 python scripts/ncu_targets.py $1 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"; tail -2 gpurun_out/ncu_$1.log
}
cap epoch bprmf_epoch_owner_kernel 1 1
cap adam_marked 'adam_sweep_marked_kernel|mark_rows_kernel' 2 2
cap spmm_cfg4_masked csr_spmm_kernel 1 1
cap eval_x64 'eval_tc_rank_kernel|eval_recheck_kernel' 2 2
ls -la gpurun_out/r2b_*.ncu-rep
