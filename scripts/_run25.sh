cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=line -k "epoch or resident or stream or host_fed or ml100k_epoch_bprmf" 2>&1 | grep -E "^FAILED|passed|failed|Error" | head -20
timeout 200 python scripts/prof_resident.py gpurun_out/prof_resident_owner.json epoch-only 2>&1 | grep -E "us_per_step|\"bpr\"|\"adam\"|barrier2|b2_arrival_skew|b2_last_arrival_to_last_pass|finite"
