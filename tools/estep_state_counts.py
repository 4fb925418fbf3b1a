"""Development: float32 SIMT E-step against the padded tensor-core E-step by state count (B = 131072, T = 512, M = 256)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine
M, B, T = 256, 131072, 512
obs = torch.from_numpy(np.random.default_rng(1).integers(0, M, size=(B, T)).astype(np.uint8)).cuda()
ob = engine.prepare_obs(obs, symbols=True)
def tm(fn, n=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
engine.TC_MIN_STATES = 1
for N in (8, 16, 32, 64, 100):
    np.random.seed(N)
    h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
    p = h._params()
    t_tc = tm(lambda: engine.estep_discrete_tc(p, ob))
    t_simt = tm(lambda: engine.estep_discrete(p, ob, torch.float32, tensor_cores=False), 1)
    print("N %3d: SIMT %.1f ms, tensor cores (padded to 128) %.1f ms" % (N, t_simt, t_tc))
