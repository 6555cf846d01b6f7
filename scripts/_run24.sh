cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=line -k "epoch or resident or stream or host_fed or ml100k_epoch_bprmf" 2>&1 | tail -25
