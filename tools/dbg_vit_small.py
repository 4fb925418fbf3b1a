"""Development: Viterbi of small batches (KHMM_VARIANT=mb1: the screened kernel from one sequence up)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from kerehmm_b200 import _lib
if os.environ.get('KHMM_VARIANT'):
    _lib.LIB_PATH = os.path.join(_lib.PKG, 'libkhmm_%s.so' % os.environ['KHMM_VARIANT'])
import kerehmm_b200 as K
from kerehmm_b200 import engine
for N in (32, 64, 128):
    np.random.seed(N)
    h = K.DiscreteHMM(N, 64, random_transitions=True, random_emissions=True)
    p = h._params()
    for B in (1, 8, 32):
        obs = torch.from_numpy(np.random.default_rng(B).integers(0, 64, size=(B, 4096)).astype(np.uint8)).cuda()
        ob = engine.prepare_obs(obs, symbols=True)
        engine.viterbi(p, ob); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): path, logp = engine.viterbi(p, ob)
        e1.record(); torch.cuda.synchronize()
        print(os.environ.get('KHMM_VARIANT'), "N", N, "B", B, "%.2f ms" % (e0.elapsed_time(e1) / 3), int(path.sum()), float(logp.sum()))
