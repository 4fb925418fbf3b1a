import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine
x = torch.randn(16384, 2048).pin_memory()
d = torch.empty_like(x, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("raw pinned H2D of 134 MB: %.3f ms = %.1f GB/s" % (dt * 1e3, x.numel() * 4 / dt / 1e9))
np.random.seed(2)
g = K.GaussianHMM(8, 1, random_transitions=True, random_emissions=True)
p = g._params()
for n in (2, 4, 8, 16):
    engine.fwdbwd_loglik_small_host(p, x, nchunks=n)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): engine.fwdbwd_loglik_small_host(p, x, nchunks=n)
    torch.cuda.synchronize(); print("nchunks", n, "%.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
