cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 > gpurun_out/bench_n8_v2.json 2> gpurun_out/bench_n8_v2.err; echo "bench rc=$?"; tail -c 800 gpurun_out/bench_n8_v2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n8_v2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e']['value'])
ex=d.get('extra',{})
for k,v in ex.items():
    print(k, json.dumps(v)[:500])
"
