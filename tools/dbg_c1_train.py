"""Development: wall time of the reference-named single-sequence training calls at the c1 shape."""
import sys, time, contextlib, io, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
np.random.seed(1)
h = K.DiscreteHMM(3, 4, random_transitions=True, random_emissions=True)
obs = np.random.default_rng(1001).integers(0, 4, size=1000)
h.do_pass(obs); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): h.do_pass(obs)
torch.cuda.synchronize(); print("do_pass(T=1000, 3x4): %.2f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
np.random.seed(1)
h = K.DiscreteHMM(3, 4, random_transitions=True, random_emissions=True)
t0 = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()):
    h.train(obs, iterations=20, auto_stop=False)
torch.cuda.synchronize(); print("train(iterations=20) = 21 passes: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
