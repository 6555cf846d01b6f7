cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/prof_sharded.py gpurun_out/r2_prof_sharded_n2.json bprmf 2>&1 | grep -A40 '^{' | head -50
