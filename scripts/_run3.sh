cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "csr_build or propagation_engine or resident or host_fed or bprmf" 2>&1 | tail -15
