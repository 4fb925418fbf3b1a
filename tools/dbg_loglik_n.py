"""Development: log-likelihood of mid-size state spaces: float32 SIMT forward kernel against the padded tensor-core recursion."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine
from oracle import hmm_oracle as O
B, T = 16384, 1024
def tm(fn, n=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
engine.TC_MIN_STATES = 1
for N in (20, 32, 64):
    np.random.seed(N)
    pi, A, mean, sd = O.random_gaussian_model(N)
    h = K.GaussianHMM(N, 1)
    h.initialProbabilities, h.transitionMatrix = pi.copy(), A.copy()
    h.emissionDistributions = [K.GaussianDistribution(1, mean=float(m), variance=float(s)) for m, s in zip(mean, sd)]
    rng = np.random.default_rng(N + 1)
    x = torch.from_numpy((mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)).cuda()
    ref = O.loglik(O.forward(pi, A, O.emissions_gaussian_1d(mean, sd, x[:16].cpu().numpy().astype(np.float64)))[1])
    l_s = h.log_likelihood_batch(x, tensor_cores=False)[0].cpu().numpy()[:16]
    l_t = h.log_likelihood_batch(x, tensor_cores=True)[0].cpu().numpy()[:16]
    print("N %3d: SIMT %.1f ms (rel err %.1e), tensor cores padded to 128: %.1f ms (rel err %.1e)" % (
        N, tm(lambda: h.log_likelihood_batch(x, tensor_cores=False), 1), np.abs(l_s / ref - 1).max(),
        tm(lambda: h.log_likelihood_batch(x, tensor_cores=True)), np.abs(l_t / ref - 1).max()))
