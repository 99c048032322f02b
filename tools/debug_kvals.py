import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mlatom_b200 import engine as eng, synthetic as S
nat=9; eq=S.equilibrium(nat); dev=lambda a: torch.as_tensor(np.ascontiguousarray(a,dtype=np.float64)).cuda()
xeq=eng.xeq(dev(eq))
for ntr in (1000, 5000, 50000):
    xyz=dev(S.geometries(eq,ntr,1))
    k=eng.build_kernel_matrix(xyz,xeq,1.0,1e-6,1e-6)
    rng=np.random.default_rng(11); rows=np.sort(rng.choice(ntr,size=48,replace=False))
    blk=eng.kernel_block(xyz[torch.as_tensor(rows).cuda()].contiguous(),xyz,xeq,1.0)
    worst=(0,0,0)
    for n,i in enumerate(rows):
        if i==0: continue
        d=(k[i,:i]-blk[n,:i]).abs()
        m=d.max().item()
        if not (m<=worst[0]): worst=(m,int(i),int(d.argmax()))
    print(ntr,"worst",worst, "nan in k rows", bool(torch.isnan(k[torch.as_tensor(rows).cuda()]).any()), "nan blk", bool(torch.isnan(blk).any()))
    i,j=worst[1],worst[2]
    print("  k",k[i,j].item(),"blk",blk[list(rows).index(i),j].item())
