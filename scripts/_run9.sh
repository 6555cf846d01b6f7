cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "exact_tensor or tcgen05 or tensor or eval" 2>&1 | tail -3
python - <<'P'
import torch, numpy as np, sys
sys.path.insert(0,'.')
from whisprrec_b200 import _lib
dev=torch.device('cuda'); g=torch.Generator(device=dev); g.manual_seed(3407)
for d in (64,128):
    nUs,nIs,Rs=200_000,1_000_000,262_144
    Ub=torch.randn((nUs,d),device=dev,generator=g)/d**0.5; Ib=torch.randn((nIs,d),device=dev,generator=g)
    us=torch.randint(0,nUs,(Rs,),device=dev,generator=g); ps=torch.randint(0,nIs,(Rs,),device=dev,generator=g)
    hp_=torch.arange(0,(nUs+1)*50,50,device=dev,dtype=torch.int64)
    hi_=torch.sort(torch.randint(0,nIs,(nUs,50),device=dev,generator=g),dim=1).values.to(torch.int32).reshape(-1).contiguous()
    ws=_lib.Workspace(dev)
    for prec in (1,2):
        fn=lambda: _lib.eval_rank_topk(Ub,Ib,us,ps,hp_,hi_,ws,precision=prec)
        fn(); torch.cuda.synchronize()
        ms=[]
        for _ in range(3):
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        print('D',d,'precision',prec,'ms',np.median(ms), 'TF(fp32-eq)', 2.0*Rs*nIs*d/(np.median(ms)*1e-3)/1e12)
P
