"""Development: fused tensor-core E-step on a ragged batch at the c4 size against the equal-length batch and the SIMT path."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine
N, M, B, T = 128, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 131072, 512
np.random.seed(4)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
rng = np.random.default_rng(1004)
obs = torch.from_numpy(rng.integers(0, M, size=(B, T)).astype(np.uint16).view(np.int16)).cuda()
lengths = rng.integers(T // 2, T + 1, size=B)
p = h._params()
def tm(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ob_eq = engine.prepare_obs(obs, symbols=True, u16_bits=True)
ob_rg = engine.prepare_obs(obs, symbols=True, lengths=lengths, u16_bits=True)
print("equal-length TC E-step %.2f ms" % tm(lambda: engine.estep_discrete_tc(p, ob_eq)))
print("ragged (lengths U[T/2, T]) TC E-step %.2f ms" % tm(lambda: engine.estep_discrete_tc(p, ob_rg)))
nb = 8192
ob_s = engine.prepare_obs(obs[:nb], symbols=True, lengths=lengths[:nb], u16_bits=True)
print("ragged SIMT E-step on %d sequences %.2f ms (x%d for the batch)" % (nb, tm(lambda: engine.estep_discrete(p, ob_s, torch.float32, tensor_cores=False), 1), B // nb))
