import os, sys, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine
N, M, B, T = 128, 1024, 16384, 512
np.random.seed(3)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
obs = torch.from_numpy(np.random.default_rng(1003).integers(0, M, size=(B, T)).astype(np.int16)).cuda()
ob = engine.prepare_obs(obs, symbols=True)
p = h._params()
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); path, logp = engine.viterbi(p, ob); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print("N=128 screen", os.environ.get("KHMM_VIT_SCREEN", "1"), "ms %.1f" % ms, "transitions/s %.3e" % (B * T * N * N / ms * 1e3), int(path.sum().item()), float(logp.sum().item()))
