import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import kerehmm_b200 as K
from kerehmm_b200 import engine
N, M = 64, 256
np.random.seed(3)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
p = h._params()
for B, T in ((16384, 256), (65536, 512), (65536, 4096)):
    obs = torch.randint(0, M, (B, T), dtype=torch.uint8, device='cuda')
    ob = engine.prepare_obs(obs, symbols=True)
    for _ in range(2): engine.viterbi(p, ob, torch.uint8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); engine.viterbi(p, ob, torch.uint8); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(B, T, "%.2f ms %.3e tr/s" % (ms, B * T * N * N / ms * 1e3))
