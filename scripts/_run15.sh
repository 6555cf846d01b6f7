cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "marked or adam or bprmf_step or large" 2>&1 | tail -6
timeout 300 python - <<'P'
import torch, time
from whisprrec_b200 import _lib
dev=torch.device('cuda')
nU,nI,d,b=10_000_000,2_000_000,128,65536
P=torch.empty((nU+nI,d),device=dev).normal_(0,0.01)
M,V,G=torch.zeros_like(P),torch.zeros_like(P),torch.zeros_like(P)
ws,loss=_lib.Workspace(dev),torch.zeros(1,device=dev)
u=torch.randint(0,nU,(b,),device=dev);p=torch.randint(0,nI,(b,),device=dev);n=torch.randint(1,nI,(b,),device=dev)
t=_lib.row_map(nU+nI,dev)
def tm(f,reps=6):
    f();torch.cuda.synchronize();r=[]
    for _ in range(reps):
        a,bb=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record();f();bb.record();torch.cuda.synchronize();r.append(a.elapsed_time(bb))
    return sorted(r)[len(r)//2]
k=[0]
def marked():
    k[0]+=1;_lib.bprmf_step(P,M,V,G,u,p,n,nU,k[0],1e-3,1e-6,loss,ws,touched=t)
def dense():
    k[0]+=1;_lib.bprmf_step(P,M,V,G,u,p,n,nU,k[0],1e-3,1e-6,loss,ws)
print('marked ms',tm(marked),'dense ms',tm(dense))
print('sweep marked only', tm(lambda:_lib.adam_l2_sweep_marked(P,M,V,G,t,1,1e-3,1e-6)), 'dense only', tm(lambda:_lib.adam_l2_sweep(P,M,V,G,1,1e-3,1e-6)))
P
