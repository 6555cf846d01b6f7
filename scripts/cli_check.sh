#!/bin/bash
# Drive the reference-compatible command line end to end on the ml-1m-shaped stand-in: one GPU, then the same run
# row-sharded over two GPUs (torchrun).  The two logs must report the same losses / dev metrics epoch by epoch.
#   bash scripts/cli_check.sh [MODEL] [EPOCHS] [NGPU]
set -e
MODEL=${1:-BPRMF}; EPOCHS=${2:-3}; NGPU=${3:-2}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
DATA=/tmp/wr_cli/data; mkdir -p /tmp/wr_cli/run && cd /tmp/wr_cli/run
python - <<PY
import os, sys
sys.path.insert(0, "$ROOT")
from whisprrec_b200.utils import synthetic
p = "$DATA/ml-1m/ml-1m.inter"
if not os.path.exists(p):
    synthetic.write_inter(synthetic.ml1m_shaped(), p)
PY
HP="--lr 1e-3 --l2 1e-6"; [ "$MODEL" = LightGCN ] && HP="--lr 5e-4 --l2 0 --gcn_layers 2"     # what the golden logs were recorded with
COMMON="--model_name $MODEL --emb_size 64 $HP --dataset ml-1m --path $DATA/ --epoch $EPOCHS --regenerate 1"
PYTHONPATH=$ROOT python -m whisprrec_b200.main $COMMON --log_file /tmp/wr_cli/one.log --model_path /tmp/wr_cli/one.pt > /tmp/wr_cli/one.out 2>&1
grep -E "Epoch|Test After" /tmp/wr_cli/one.out | sed 's/\[[0-9. ]*s\]//g' > /tmp/wr_cli/one.txt
# the unmodified reference's log on the same file, where one was recorded (tests/golden/ml1m_shaped_reference_log*.txt)
GOLD=$ROOT/tests/golden/ml1m_shaped_reference_log.txt
[ "$MODEL" = LightGCN ] && GOLD=$ROOT/tests/golden/ml1m_shaped_reference_log_lightgcn.txt
if [ "$EPOCHS" = 3 ] && [ -f "$GOLD" ]; then
  if diff -q <(grep -v '^#' "$GOLD" | tr -s ' ' | sed 's/ *$//') <(tr -s ' ' < /tmp/wr_cli/one.txt | sed 's/ *$//') > /dev/null; then
    echo "cli_check ok: equals the reference's log"; else echo "cli_check: differs from the reference's log"; RC=1; fi
fi
if [ "$NGPU" -gt 1 ]; then
  PYTHONPATH=$ROOT python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29577 \
      -m whisprrec_b200.main $COMMON --log_file /tmp/wr_cli/multi.log --model_path /tmp/wr_cli/multi.pt > /tmp/wr_cli/multi.out 2>&1
  grep -E "Epoch|Test After" /tmp/wr_cli/multi.out | sed 's/\[[0-9. ]*s\]//g' | sort -u > /tmp/wr_cli/multi.txt
  sort -u /tmp/wr_cli/one.txt > /tmp/wr_cli/one_sorted.txt
  echo "--- one GPU";  cat /tmp/wr_cli/one.txt
  echo "--- $NGPU GPUs"; cat /tmp/wr_cli/multi.txt
  if diff -q /tmp/wr_cli/one_sorted.txt /tmp/wr_cli/multi.txt > /dev/null; then echo "cli_check ok: identical log lines"; else echo "cli_check: logs differ"; diff /tmp/wr_cli/one_sorted.txt /tmp/wr_cli/multi.txt | head; RC=1; fi
else
  cat /tmp/wr_cli/one.txt; echo "cli_check ok (one GPU)"
fi
exit ${RC:-0}
