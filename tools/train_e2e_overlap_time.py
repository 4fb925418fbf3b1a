"""Development: end-to-end time of one Baum-Welch iteration through DiscreteHMM.train_batch on a page-locked config-4 batch,
with the overlapped upload (three row blocks on a side stream, E-step block by block) and without it."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine
N, M, B, T = 128, 1024, 131072, 512
np.random.seed(4)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
host = torch.from_numpy(np.random.default_rng(1004).integers(0, M, size=(B, T)).astype(np.uint16).view(np.int16)).pin_memory()
arr = host.numpy().view(np.uint16)
print("pinned view:", torch.from_numpy(arr.view(np.int16)).is_pinned())
default_min = engine.OVERLAP_MIN_BYTES
for name, min_bytes in (("overlapped", default_min), ("plain", 1 << 60), ("overlapped", default_min), ("plain", 1 << 60)):
    engine.OVERLAP_MIN_BYTES = min_bytes
    for _ in range(2):
        h.train_batch(arr, iterations=1)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        h.train_batch(arr, iterations=1)
        torch.cuda.synchronize()
        ts.append(round((time.perf_counter() - t0) * 1e3, 2))
    print(name, "ms per call", ts, flush=True)
