import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine
from oracle import hmm_oracle as O
N, B, T = 1024, 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 64
np.random.seed(5)
pi, A, mean, sd = O.random_gaussian_model(N)
h = K.GaussianHMM(N, 1)
h.initialProbabilities, h.transitionMatrix = pi.copy(), A.copy()
h.emissionDistributions = [K.GaussianDistribution(1, mean=float(m), variance=float(s)) for m, s in zip(mean, sd)]
rng = np.random.default_rng(6)
x = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
xd = torch.from_numpy(x).cuda()
ref = O.loglik(O.forward(pi, A, O.emissions_gaussian_1d(mean, sd, x[:64].astype(np.float64)))[1])
for mode in (True, "bf16"):
    ll, llb = h.log_likelihood_batch(xd, tensor_cores=mode)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ll, llb = h.log_likelihood_batch(xd, tensor_cores=mode)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    l, lb = ll.cpu().numpy()[:64], llb.cpu().numpy()[:64]
    print(mode, "ms %.2f" % (dt * 1e3), "fwd max rel err vs f64 oracle %.3e" % np.max(np.abs(l - ref) / np.abs(ref)),
          "bwd %.3e" % np.max(np.abs(lb - ref) / np.abs(ref)), "mean signed rel %.3e" % np.mean((l - ref) / np.abs(ref)))
