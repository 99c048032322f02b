import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mlatom_b200 import engine as eng, synthetic as S, _lib
nat=9; eq=S.equilibrium(nat); dev=lambda a: torch.as_tensor(np.ascontiguousarray(a,dtype=np.float64)).cuda()
xeq=eng.xeq(dev(eq)); ntr=5000
xyz=dev(S.geometries(eq,ntr,1))
for rep in range(3):
    k=eng.build_kernel_matrix(xyz,xeq,1.0,1e-6,1e-6)
    out=(ctypes.c_ulonglong*8)()
    lib=_lib.load(); lib.kreg_debug_bias.argtypes=[ctypes.c_void_p]; lib.kreg_debug_bias(out)
    print("early!=truth",out[0],"late!=truth",out[1],"early!=late",out[2],"checks",out[3],"early bad at step0",out[4],"early bad at step>=NST",out[5],"wrong value is another tile's",out[6])
