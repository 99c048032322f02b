import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mlatom_b200 import engine as eng, synthetic as S
nat=9; eq=S.equilibrium(nat); dev=lambda a: torch.as_tensor(np.ascontiguousarray(a,dtype=np.float64)).cuda()
xeq=eng.xeq(dev(eq))
for ntr in ([int(x) for x in sys.argv[1:]] or [1000, 2500, 5000]):
    xyz=dev(S.geometries(eq,ntr,1))
    X=eng.descriptors(xyz,xeq)
    ref=torch.exp(-torch.cdist(X,X)**2/2.0)
    for rep in range(2):
        k=torch.full((ntr,ntr),float('nan'),dtype=torch.float64,device='cuda')
        eng.build_kernel_matrix(xyz,xeq,1.0,1e-6,1e-6,out=k)
        d=(torch.tril(k,-1)-torch.tril(ref,-1)).abs()
        bad=torch.nonzero(~(d<=1e-10))
        print(ntr,"rep",rep,"bad",bad.shape[0])
        if bad.shape[0]:
            b=bad.cpu().numpy()
            print("  first",b[:6].tolist(),"vals",[k[i,j].item() for i,j in b[:3]])
            ti=b[:,0]//128; tj=b[:,1]//64
            tiles=set(zip(ti.tolist(),tj.tolist()))
            print("  tiles",len(tiles),sorted(tiles)[:10],"rows%128 uniq",np.unique(b[:,0]%128)[:20],"cols%64 uniq",np.unique(b[:,1]%64)[:20])
            print("  tj%4 hist",np.bincount(np.array([t[1] for t in tiles])%4,minlength=4).tolist())
            print("  rows%128 full",np.unique(b[:,0]%128).tolist())
            print("  cols%64 full",np.unique(b[:,1]%64).tolist())
            t0=sorted(tiles)[0]; sel=b[(ti==t0[0])&(tj==t0[1])]
            print("  tile",t0,"nbad",len(sel),"rows",np.unique(sel[:,0]%128).tolist(),"cols",np.unique(sel[:,1]%64).tolist())
            import collections
            print("  bad per tile hist",sorted(collections.Counter(collections.Counter(zip(ti.tolist(),tj.tolist())).values()).items()))
