"""Development: accuracy (counts and re-estimated parameters against the float64 oracle) and speed of the fused tensor-core
E-step with tf32 operands and with split bf16 operands (three MMAs)."""
import ctypes, sys, os, time, numpy as np, torch
import multiprocessing as mp
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine, _lib
from oracle import hmm_oracle as O


def _stats(args):
    pi, A, Bm, obs = args
    return O.batch_stats_discrete(pi, A, Bm, obs)


def oracle_stats(pi, A, Bm, obs, procs=16):
    chunks = [(pi, A, Bm, c) for c in np.array_split(obs, procs) if len(c)]
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        parts = pool.map(_stats, chunks)
    return {k: sum(p_[k] for p_ in parts) for k in parts[0]}


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


N, M, B, T = 128, 1024, int(os.environ.get("BACC", "4096")), 512
np.random.seed(4)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
pi, A, Bm = h.initialProbabilities.copy(), h.transitionMatrix.copy(), h.emission_matrix()
obs = np.random.default_rng(1004).integers(0, M, size=(B, T)).astype(np.int64)
t0 = time.time()
st = oracle_stats(pi, A, Bm, obs)
rpi, rA, rB = O.mstep_discrete(st)
print("oracle: %d sequences in %.1f s" % (B, time.time() - t0))
ob = engine.prepare_obs(obs, symbols=True, alphabet=M)
p = h._params()
for name, prec in (("tf32", engine.TC_TF32), ("bf16x3", engine.TC_BF16X3)):
    c = engine.estep_discrete_tc(p, ob, precision=prec)
    hc = c.host()
    npi, nA, nB = engine.mstep_discrete(p, c)
    print("%-7s counts: pi %.2e  gam %.2e  emis_den %.2e  emis_num %.2e (median %.2e)  xi_norm %.2e  loglik %.2e" % (
        name, rel(hc["pi"], st["pi"]), rel(hc["gam"], st["gam"]), rel(hc["emis_den"], st["emis_den"]),
        rel(hc["emis_num"], st["emis_num"]),
        float(np.median(np.abs(hc["emis_num"] - st["emis_num"]) / st["emis_num"])),
        rel((hc["xi_raw"] * A) / (hc["xi_raw"] * A).sum(axis=1, keepdims=True), st["xi"] / st["xi"].sum(axis=1, keepdims=True)),
        rel(hc["loglik"], st["loglik"])))
    print("%-7s re-estimated parameters vs oracle M-step: pi %.2e  A %.2e  B %.2e" % (name, rel(npi, rpi), rel(nA, rA), rel(nB, rB)))

# ---- speed at config-4 size
Bf = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
obs = np.random.default_rng(1004).integers(0, M, size=(Bf, T)).astype(np.uint16)
ob = engine.prepare_obs(torch.from_numpy(obs.view(np.int16)).cuda(), symbols=True, u16_bits=True)
_lib.call("khmm_estep_tc_timing", 1)
f, b = ctypes.c_float(), ctypes.c_float()
for name, prec in (("tf32", engine.TC_TF32), ("bf16x3", engine.TC_BF16X3)):
    for _ in range(2):
        engine.estep_discrete_tc(p, ob, precision=prec)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        engine.estep_discrete_tc(p, ob, precision=prec)
        _lib.call("khmm_estep_tc_kernel_ms", 0, ctypes.byref(f), ctypes.byref(b))
        ts.append((round(f.value, 2), round(b.value, 2)))
    print("%-7s forward/backward ms at B=%d: %s" % (name, Bf, ts))
