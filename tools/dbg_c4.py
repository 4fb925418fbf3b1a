import sys, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine
from oracle import hmm_oracle as O
N, M, B, T = 128, 1024, int(sys.argv[1]), int(sys.argv[2])
np.random.seed(4)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
obs = np.random.default_rng(1004).integers(0, M, size=(B, T)).astype(np.uint16)
ob = engine.prepare_obs(torch.from_numpy(obs.view(np.int16)).cuda(), symbols=True, u16_bits=True)
for it in range(3):
    p = h._params()
    c_tc, W, sc = engine.estep_discrete_tc(p, ob, stash_out=True)
    c_ref = engine.estep_discrete(p, ob, torch.float32, tensor_cores=False)
    a, b = c_tc.host(), c_ref.host()
    print("iter", it, "stash finite", bool(torch.isfinite(W).all()), "scale finite", bool(torch.isfinite(sc).all()), "scale min", float(sc.min()))
    for k in ("pi", "gam", "emis_den", "emis_num", "xi_raw", "loglik", "nseq", "flags"):
        x, y = np.asarray(a[k], dtype=np.float64), np.asarray(b[k], dtype=np.float64)
        bad = ~np.isfinite(x)
        rel = np.abs(x - y) / np.maximum(np.abs(y), 1e-300)
        print("   %-9s nonfinite %d  max rel %.3e  median rel %.3e  sum tc %.9g ref %.9g" % (k, bad.sum(), np.nanmax(rel), np.nanmedian(rel), np.nansum(x), np.nansum(y)))
    try:
        h._assign(*engine.mstep_discrete(p, c_tc))
        print("   mstep ok")
    except Exception as e:
        print("   mstep FAILED:", e)
        h._assign(*engine.mstep_discrete(p, c_ref))
