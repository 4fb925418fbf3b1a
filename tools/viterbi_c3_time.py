"""Development: time Viterbi at the c3 shape (KHMM_VARIANT=<name> picks kerehmm_b200/libkhmm_<name>.so)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from kerehmm_b200 import _lib
if os.environ.get('KHMM_VARIANT'):
    _lib.LIB_PATH = os.path.join(_lib.PKG, 'libkhmm_%s.so' % os.environ['KHMM_VARIANT'])
import kerehmm_b200 as K
from kerehmm_b200 import engine
N, M = 64, 256
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
np.random.seed(3)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
obs = torch.from_numpy(np.random.default_rng(1003).integers(0, M, size=(B, T)).astype(np.uint8)).cuda()
ob = engine.prepare_obs(obs, symbols=True)
p = h._params()
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    path, logp = engine.viterbi(p, ob)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print("variant", os.environ.get('KHMM_VARIANT'), "screen", os.environ.get("KHMM_VIT_SCREEN", "1"), "ms %.1f" % ms,
      "transitions/s %.3e" % (B * T * N * N / ms * 1e3), "checksum", int(path.sum().item()), float(logp.sum().item()))
