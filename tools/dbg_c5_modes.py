"""Development: time and check the whole-recursion kernel generations on the c5 shape.
usage: python tools/dbg_c5_modes.py [T] [modes e.g. 1,2,3] ; KHMM_RC_MODE is read by the library on every call."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
from kerehmm_b200 import _lib
if os.environ.get('KHMM_VARIANT'):
    _lib.LIB_PATH = os.path.join(_lib.PKG, 'libkhmm_%s.so' % os.environ['KHMM_VARIANT'])
import kerehmm_b200 as K
from oracle import hmm_oracle as O
N, B = 1024, 4096
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
modes = [int(m) for m in (sys.argv[2] if len(sys.argv) > 2 else "1,2,3").split(",")]
np.random.seed(5)
pi, A, mean, sd = O.random_gaussian_model(N)
h = K.GaussianHMM(N, 1)
h.initialProbabilities, h.transitionMatrix = pi.copy(), A.copy()
h.emissionDistributions = [K.GaussianDistribution(1, mean=float(m), variance=float(s)) for m, s in zip(mean, sd)]
rng = np.random.default_rng(6)
x = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
xd = torch.from_numpy(x).cuda()
nref = 16
ref = O.loglik(O.forward(pi, A, O.emissions_gaussian_1d(mean, sd, x[:nref].astype(np.float64)))[1])
for mode in modes:
    os.environ["KHMM_RC_MODE"] = str(mode)
    for prec in (True, "bf16"):
        ll, llb = h.log_likelihood_batch(xd, tensor_cores=prec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ll, llb = h.log_likelihood_batch(xd, tensor_cores=prec)
        e1.record(); torch.cuda.synchronize()
        l, lb = ll.cpu().numpy(), llb.cpu().numpy()
        print("mode", mode, "bf16" if prec == "bf16" else "tf32", "ms %.2f" % (e0.elapsed_time(e1) / 3),
              "fwd rel err %.2e" % np.max(np.abs(l[:nref] - ref) / np.abs(ref)),
              "bwd %.2e" % np.max(np.abs(lb[:nref] - ref) / np.abs(ref)),
              "fwd-bwd all rows %.2e" % np.max(np.abs(l - lb) / np.abs(l)), flush=True)
