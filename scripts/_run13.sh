cd $GRAFT_REPO_ROOT
for v in 0 1 2 3 11; do
WR_TC_VARIANT=$v python - <<'P'
import torch, numpy as np, sys, os
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
    fn=lambda: _lib.eval_rank_topk(Ub,Ib,us,ps,hp_,hi_,ws,precision=1)
    fn(); torch.cuda.synchronize()
    ms=[]
    for _ in range(3):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    print('variant',os.environ['WR_TC_VARIANT'],'D',d,'ms %.2f'%np.median(ms), 'TF %.0f'%(2.0*Rs*nIs*d/(np.median(ms)*1e-3)/1e12))
P
done
