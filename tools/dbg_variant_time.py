"""Development: time the tensor-core E-step kernels of a libkhmm variant (KHMM_VARIANT=<name>) at full c4 size."""
import ctypes, os, sys, numpy as np, torch
sys.path.insert(0, '.')
from kerehmm_b200 import _lib
if os.environ.get('KHMM_VARIANT'):
    _lib.LIB_PATH = os.path.join(_lib.PKG, 'libkhmm_%s.so' % os.environ['KHMM_VARIANT'])
import kerehmm_b200 as K
from kerehmm_b200 import engine
N, M, B, T = 128, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 131072, 512
np.random.seed(4)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
obs = np.random.default_rng(1004).integers(0, M, size=(B, T)).astype(np.uint16)
ob = engine.prepare_obs(torch.from_numpy(obs.view(np.int16)).cuda(), symbols=True, u16_bits=True)
p = h._params()
res = []
for it in range(4):
    engine.estep_discrete_tc(p, ob)
    f, b = ctypes.c_float(), ctypes.c_float()
    _lib.call("khmm_estep_tc_last_kernel_ms", ctypes.byref(f), ctypes.byref(b))
    res.append((f.value, b.value))
print("variant", os.environ.get('KHMM_VARIANT'), "fwd ms %.3f bwd ms %.3f" % tuple(np.mean(res[1:], axis=0)))
