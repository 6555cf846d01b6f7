cd $GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 scripts/prof_push_overlap.py gpurun_out/prof_push_overlap_n8_v3.json 1.0 4 2>&1 | grep -E "spmm|dma"
WR_SPMM_PUSH=store timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 scripts/prof_sharded.py gpurun_out/prof_sharded_n8_v8_store_push.json lightgcn 2>&1 | grep -E "step|sum_csr|sum_push|sum_peer"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 scripts/prof_sharded.py gpurun_out/prof_sharded_n8_v8_dma_push.json lightgcn 2>&1 | grep -E "step|sum_csr|sum_push|sum_peer"
