import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine, _lib
N, M, B, T = 128, 1024, int(sys.argv[1]), 512
np.random.seed(4)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
obs = np.random.default_rng(1004).integers(0, M, size=(B, T)).astype(np.uint16)
ob = engine.prepare_obs(torch.from_numpy(obs.view(np.int16)).cuda(), symbols=True, u16_bits=True)
def sync(): torch.cuda.synchronize()
for it in range(4):
    sync(); t0 = time.perf_counter()
    p = h._params(); t1 = time.perf_counter()
    counts = engine.estep_discrete(p, ob, torch.float32, tensor_cores=True); t2 = time.perf_counter()
    sync(); t3 = time.perf_counter()
    res = engine.mstep_discrete(p, counts); sync(); t4 = time.perf_counter()
    h._assign(*res); t5 = time.perf_counter()
    print("iter %d: params %.2f ms, estep launch %.2f ms, estep wait %.2f ms, mstep %.2f ms, assign %.2f ms, total %.2f ms" % (
        it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3, (t5-t0)*1e3))
