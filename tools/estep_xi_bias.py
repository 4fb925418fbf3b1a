"""Development: systematic relative difference of the tensor-core E-step statistics against the float32 SIMT path by N."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
M, B, T = 24, 148 * 128 + 50, 8
for N in (65, 96, 100, 112, 128):
    np.random.seed(90 + N)
    h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
    obs = np.random.default_rng(91 + N).integers(0, M, size=(B, T)).astype(np.uint8)
    ref = h.estep_batch(obs, tensor_cores=False).host()
    c = h.estep_batch(obs, tensor_cores=True).host()
    out = []
    for key in ("pi", "gam", "emis_den", "emis_num", "xi_raw"):
        rel = (c[key] - ref[key]) / np.abs(ref[key]).clip(1e-30)
        out.append("%s med %+.1e max %.1e" % (key, np.median(rel), np.abs(rel).max()))
    print("N", N, " | ".join(out))
