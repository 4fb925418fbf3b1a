import sys, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from oracle import hmm_oracle as O
N, B, T = 512, 4864, 6
np.random.seed(5)
pi, A, mean, sd = O.random_gaussian_model(N)
h = K.GaussianHMM(N, 1)
h.initialProbabilities, h.transitionMatrix = pi.copy(), A.copy()
h.emissionDistributions = [K.GaussianDistribution(1, mean=float(m), variance=float(s)) for m, s in zip(mean, sd)]
rng = np.random.default_rng(6)
x = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
ll, llb = h.log_likelihood_batch(torch.from_numpy(x).cuda())
ll2, _ = h.log_likelihood_batch(torch.from_numpy(x).cuda(), tensor_cores=False)
ll, llb, ll2 = ll.cpu().numpy(), llb.cpu().numpy(), ll2.cpu().numpy()
print("max rel fwd vs simt", np.max(np.abs(ll - ll2) / np.abs(ll2)), "bwd vs simt", np.max(np.abs(llb - ll2) / np.abs(ll2)), "nan", np.isnan(ll).sum(), np.isnan(llb).sum())
