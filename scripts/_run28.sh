cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 > gpurun_out/bench_n2_v1.json 2> gpurun_out/bench_n2_v1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_n2_v1.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n2_v1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e']['value'])
ex=d.get('extra',{})
for k,v in ex.items():
    print(k, json.dumps(v)[:300])
"
