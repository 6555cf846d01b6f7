cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=line -k "epoch or resident or ml100k_epoch_bprmf" 2>&1 | grep -E "^FAILED|passed|failed|Error" | head -20
timeout 200 python scripts/prof_owner.py gpurun_out/prof_owner_v6.json 2>&1 | tail -19
