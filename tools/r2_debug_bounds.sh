# Memory-safety substitute (compute-sanitizer is closed on this pool): the GPU parity suite against the bounds-checking
# build of the library; the in-kernel failure counter must stay at zero.
set -x
cd $GRAFT_REPO_ROOT
export KREG_LIB=$PWD/mlatom_b200/libkreg_b200_dbg.so
timeout 2400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large_molecules.py tests/test_gpu_other_kernels.py tests/test_gpu_api.py tests/test_gpu_perminv.py -m gpu -q -p no:cacheprovider > gpurun_out/r02_debug_bounds_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_debug_bounds_tests.log
tail -5 gpurun_out/r02_debug_bounds_tests.log
python - <<'PY' | tee gpurun_out/r02_debug_bounds.txt
import ctypes, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from mlatom_b200 import _lib, engine as eng, synthetic as S
lib = _lib.load()
# one more pass in THIS process (the counter lives in the library instance): ragged shapes through every assembly kernel
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()
for nat, ntr in ((9, 333), (5, 129), (12, 77), (21, 45), (26, 23), (40, 9)):
    eq = S.equilibrium(nat); xyz = dev(S.geometries(eq, ntr, 1)); xeq = eng.xeq(dev(eq))
    for grads in (False, True):
        m = ntr + (3 * nat * ntr if grads else 0)
        eng.build_kernel_matrix(xyz, xeq, 1.3, 1e-6, 1e-6, use_gradients=grads)
        eng.build_kernel_matrix(xyz, xeq, 1.3, 1e-6, 1e-6, use_gradients=grads, fill=_lib.FILL_FULL)
        k = torch.zeros((m, m), dtype=torch.float64, device="cuda")
        cuts = [0, min(m, m // 3 + 1), min(m, 2 * m // 3 + 5), m]
        for r0, r1 in zip(cuts[:-1], cuts[1:]):
            eng.build_kernel_matrix(xyz, xeq, 1.3, 1e-6, 1e-6, use_gradients=grads, rows=(r0, r1), out=k)
    a = torch.randn(ntr + 3 * nat * ntr, dtype=torch.float64, device="cuda")
    model = eng.PackedModel.from_training_set(xyz, xeq, a, 1.3, 0.0, ntr, 3 * nat * ntr)
    model.predict(dev(S.geometries(eq, 301, 2)))
    for kind in ("exponential", "matern", "laplacian"):
        eng.kernel_block_ex(kind, xyz, dev(S.geometries(eq, 57, 3)), xeq, 1.1)
    from mlatom_b200 import kernels, perminv
    for kind in ("matern", "exponential"):
        kernels.RadialModel(kind, 2, xyz, xeq, a[:ntr].contiguous(), 1.1, 0.0).predict(dev(S.geometries(eq, 301, 2)))
    perms = perminv.Permutations(perminv.atom_maps(nat, [[0, 1], [2, 3, 4]] if nat >= 5 else [[0, 1]]), torch.device("cuda"))
    perminv.kernel_block(xyz, dev(S.geometries(eq, 57, 3)), xeq, perms, 1.1)
    perminv.PermInvModel(xyz, xeq, a[:ntr].contiguous(), 1.1, 0.0, perms).predict(dev(S.geometries(eq, 301, 2)))
code, line = ctypes.c_int(0), ctypes.c_int(0)
n = lib.kreg_debug_bounds_failures(ctypes.byref(code), ctypes.byref(line))
print(f"kreg_debug_bounds_failures = {n} (first code {code.value}, line {line.value}); library = {_lib.LIB_PATH}")
assert n == 0, "bounds / alignment check failed inside a kernel"
PY
