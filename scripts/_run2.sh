cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "resident or host_fed or epoch_call or bprmf" 2>&1 | tail -5
timeout 500 python scripts/prof_resident.py gpurun_out/r2_prof_resident_v2.json 2>&1 | tail -130
