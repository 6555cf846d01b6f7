cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "epoch or resident or stream or host_fed or ml100k_epoch_bprmf" 2>&1 | tail -5
timeout 200 python scripts/prof_resident.py gpurun_out/prof_resident_owner.json epoch-only 2>&1 | tail -32
WR_EPOCH_KERNEL=two_barrier timeout 200 python scripts/prof_resident.py gpurun_out/prof_resident_two.json epoch-only 2>&1 | grep -E "us_per_step|ms_median"
