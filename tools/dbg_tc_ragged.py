import sys, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from oracle import hmm_oracle as O
for N in (1000, 1024, 896):
    B, T = 64, 24
    rng = np.random.default_rng(600 + N)
    lengths = rng.integers(1, T + 1, size=B)
    np.random.seed(601 + N)
    pi, A, mean, sd = O.random_gaussian_model(N)
    h = K.GaussianHMM(N, 1)
    h.initialProbabilities, h.transitionMatrix = pi.copy(), A.copy()
    h.emissionDistributions = [K.GaussianDistribution(1, mean=float(m), variance=float(s)) for m, s in zip(mean, sd)]
    x = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)
    E = O.emissions_gaussian_1d(mean, sd, x.astype(np.float64))
    ref_r = np.array([O.loglik(O.forward(pi, A, E[b, :lengths[b]])[1]) for b in range(B)])
    ll, llb = h.log_likelihood_batch(torch.from_numpy(x).cuda(), lengths=lengths, tensor_cores=True)
    ll, llb = ll.cpu().numpy(), llb.cpu().numpy()
    npad = T - lengths
    print("N", N, "fwd err max %.2e" % np.abs(ll - ref_r).max(), "bwd err max %.2e" % np.abs(llb - ref_r).max())
    o = np.argsort(npad)
    print("  npad:", npad[o][::6]); print("  bwd err:", np.round((llb - ref_r)[o][::6], 5)); print("  fwd err:", np.round((ll - ref_r)[o][::6], 5))
