cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 scripts/prof_sharded.py gpurun_out/prof_sharded_n8_v6.json both 2>&1 | tail -75
