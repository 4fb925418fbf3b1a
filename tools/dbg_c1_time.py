"""Development: per-call device time of the single-sequence generic kernels at the c1 shape (KHMM_VARIANT picks a library)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from kerehmm_b200 import _lib
if os.environ.get('KHMM_VARIANT'):
    _lib.LIB_PATH = os.path.join(_lib.PKG, 'libkhmm_%s.so' % os.environ['KHMM_VARIANT'])
import kerehmm_b200 as K
from kerehmm_b200 import engine
N, M, T = 3, 4, 1000
np.random.seed(1)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
obs = torch.from_numpy(np.random.default_rng(1001).integers(0, M, size=(1, T))).cuda()
ob = engine.prepare_obs(obs, symbols=True)
p = h._params()
def tm(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
_, scale, _ = engine.forward(p, ob, torch.float64)
print("variant", os.environ.get('KHMM_VARIANT'), "forward %.3f ms" % tm(lambda: engine.forward(p, ob, torch.float64)),
      "backward %.3f ms" % tm(lambda: engine.backward(p, ob, torch.float64, scale)),
      "viterbi %.3f ms" % tm(lambda: engine.viterbi(p, ob)),
      "DeviceModel %.3f ms" % tm(lambda: engine.DeviceModel(p, torch.float64)))
