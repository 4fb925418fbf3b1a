"""Development: clock64 timeline of the whole-recursion tensor-core kernel (block 0, steps 8..11, both lanes)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine, _lib
from oracle import hmm_oracle as O
N, B, T = 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 4096, 64
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
np.random.seed(5)
pi, A, mean, sd = O.random_gaussian_model(N)
h = K.GaussianHMM(N, 1)
h.initialProbabilities, h.transitionMatrix = pi.copy(), A.copy()
h.emissionDistributions = [K.GaussianDistribution(1, mean=float(m), variance=float(s)) for m, s in zip(mean, sd)]
rng = np.random.default_rng(6)
xd = torch.from_numpy((mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)).cuda()
h.log_likelihood_batch(xd, tensor_cores=mode if mode == "bf16" else True)
tr = torch.zeros(4 * 2 * 16, dtype=torch.int64, device="cuda")
_lib.call("khmm_debug_rc_trace", tr.data_ptr())
h.log_likelihood_batch(xd, tensor_cores=mode if mode == "bf16" else True)
torch.cuda.synchronize()
_lib.call("khmm_debug_rc_trace", None)
a = tr.cpu().numpy().reshape(8, 16)
names = {0: "tma:job", 1: "tma:deps ok", 4: "mma:acc free", 5: "mma:kb1", 6: "mma:issued", 8: "epi:job", 9: "epi:deps ok", 10: "epi:acc full", 11: "epi:published", 12: "epi:ld0", 13: "epi:chunk0", 14: "epi:loop end", 15: "epi:fenced"}
base = a[0, 0]
for j in range(8):
    ev = sorted((int(a[j, k] - base), names[k]) for k in names if a[j, k])
    print("job", j + 16, " | ".join("%s @%d" % (n, c) for c, n in ev))
