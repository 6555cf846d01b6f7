cd $GRAFT_REPO_ROOT
timeout 900 python bench.py > gpurun_out/bench_n1_v3.json 2> gpurun_out/bench_n1_v3.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_n1_v3.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/reference_arm_v2.json 2> gpurun_out/reference_arm_v2.err; echo "ref rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1_v3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e']['ms_per_step'])
print(d['roofline']['frac'], d['roofline'].get('frac_dram'), d['roofline']['resident_epoch']['us_per_step'])
print(d['clocks'])
ex=d.get('extra',{})
for k,v in ex.items():
    s=json.dumps(v)
    print(k, s[:400])
"
