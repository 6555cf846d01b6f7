cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 scripts/prof_push_overlap.py gpurun_out/prof_push_overlap_n8_v2.json 1.0 4 2>&1 | tail -18
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 scripts/prof_sharded.py gpurun_out/prof_sharded_n8_v7.json lightgcn 2>&1 | grep -E "step|sum_csr|sum_push|sum_peer"
