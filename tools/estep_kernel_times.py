"""Development: forward / backward kernel times of the fused tensor-core E-step at config-4 shape, both operand precisions.

    [KHMM_VARIANT=<name of a tools/build_variant.sh library>] [KHMM_VER=1|3] python tools/estep_kernel_times.py [sequences]"""
import ctypes, sys, os, numpy as np, torch
sys.path.insert(0, '.')
from kerehmm_b200 import _lib
if os.environ.get('KHMM_VARIANT'):
    _lib.LIB_PATH = os.path.join(_lib.PKG, 'libkhmm_%s.so' % os.environ['KHMM_VARIANT'])
import kerehmm_b200 as K
from kerehmm_b200 import engine
N, M, B, T = 128, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 131072, 512
ver = int(os.environ.get('KHMM_VER', '3'))
np.random.seed(4)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
obs = np.random.default_rng(1004).integers(0, M, size=(B, T)).astype(np.uint16)
ob = engine.prepare_obs(torch.from_numpy(obs.view(np.int16)).cuda(), symbols=True, u16_bits=True)
p = h._params()
_lib.call("khmm_estep_tc_timing", 1)
_lib.call("khmm_debug_bw_backward_version", ver)
for pname, prec in (("tf32", engine.TC_TF32), ("bf16x3", engine.TC_BF16X3)):
    for _ in range(2):
        engine.estep_discrete_tc(p, ob, precision=prec)
    torch.cuda.synchronize()
    f, b = ctypes.c_float(), ctypes.c_float()
    ts = []
    for _ in range(3):
        engine.estep_discrete_tc(p, ob, precision=prec)
        _lib.call("khmm_estep_tc_kernel_ms", 0, ctypes.byref(f), ctypes.byref(b))
        ts.append((round(f.value, 2), round(b.value, 2)))
    print("variant", os.environ.get('KHMM_VARIANT'), "version", ver, pname, "forward/backward ms", ts, flush=True)
