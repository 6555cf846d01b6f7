cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marked or row_map or propagation or lightgcn or spmm" 2>&1 | tail -4
timeout 900 python - <<'P'
import argparse, json, torch
from scripts.bench_lightgcn_scale import run
cfg = argparse.Namespace(users=10_000_000, items=2_000_000, edges=500_000_000, dim=128, layers=3, batch=65536, steps=3, warmup=1)
print(json.dumps(run(cfg, 0, 1, torch.device('cuda'))))
P
