cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 scripts/prof_push_overlap.py gpurun_out/prof_push_overlap_n8.json 1.0 4 2>&1 | tail -20
