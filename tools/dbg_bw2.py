"""Development: the transposed backward kernel (v2) against the round-1 kernel (v1) and the float64 oracle on small shapes,
then kernel times at config-4 size and a timeline."""
import ctypes, sys, os, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine, _lib
from oracle import hmm_oracle as O

L = _lib.lib()
FIELDS = ("pi", "gam", "emis_den", "emis_num", "xi_raw", "loglik", "nseq", "flags")


def run(version, p, ob):
    _lib.call("khmm_debug_bw_backward_version", version)
    c = engine.estep_discrete_tc(p, ob).host()
    torch.cuda.synchronize()
    return c


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-12 * max(1.0, np.abs(b).max()))
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


SHAPES = [(128, 16, 128, 2, False), (128, 16, 128, 3, False), (128, 16, 130, 8, False),
          (128, 48, 300, 33, True), (128, 1024, 1000, 20, False), (128, 16, 64, 1, False)]
if os.environ.get("QUICK"):
    SHAPES = SHAPES[3:4]
for (N, M, B, T, ragged) in SHAPES:
    np.random.seed(N + M + B + T)
    h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
    pi, A, Bm = h.initialProbabilities.copy(), h.transitionMatrix.copy(), h.emission_matrix()
    rng = np.random.default_rng(B * T)
    obs = rng.integers(0, M, size=(B, T))
    lengths = rng.integers(1, T + 1, size=B) if ragged else None
    ob = engine.prepare_obs(obs, symbols=True, lengths=lengths, alphabet=M)
    p = h._params()
    try:
        c1 = run(1, p, ob)
        c2 = run(2, p, ob)
    except Exception as e:
        print("shape", (N, M, B, T, ragged), "FAILED:", e)
        break
    if lengths is None:
        st = O.batch_stats_discrete(pi, A, Bm, obs)
    else:
        st = None
        for b in range(B):
            s1 = O.batch_stats_discrete(pi, A, Bm, obs[b:b + 1, :lengths[b]])
            st = s1 if st is None else {k: st[k] + s1[k] for k in st}
    ref = dict(pi=st["pi"], gam=st["gam"], emis_den=st["emis_den"], emis_num=st["emis_num"],
               xi_raw=st["xi"] / np.where(A > 0, A, 1.0), loglik=st["loglik"], nseq=st["nseq"], flags=0.0)
    print("shape N=%d M=%d B=%d T=%d ragged=%s" % (N, M, B, T, ragged))
    for f in FIELDS:
        print("   %-9s v2 vs v1 %.3e   v1 vs oracle %.3e   v2 vs oracle %.3e   finite %s" % (
            f, rel(c2[f], c1[f]), rel(c1[f], ref[f]), rel(c2[f], ref[f]), bool(np.isfinite(np.asarray(c2[f])).all())))

# ---- timing at config-4 size
N, M, B, T = 128, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 131072, 512
np.random.seed(4)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
obs = np.random.default_rng(1004).integers(0, M, size=(B, T)).astype(np.uint16)
ob = engine.prepare_obs(torch.from_numpy(obs.view(np.int16)).cuda(), symbols=True, u16_bits=True)
p = h._params()
_lib.call("khmm_estep_tc_timing", 1)
res = {}
for version in (1, 2):
    _lib.call("khmm_debug_bw_backward_version", version)
    for _ in range(2):
        c = engine.estep_discrete_tc(p, ob)
    torch.cuda.synchronize()
    f, b = ctypes.c_float(), ctypes.c_float()
    ts = []
    for _ in range(3):
        c = engine.estep_discrete_tc(p, ob)
        _lib.call("khmm_estep_tc_kernel_ms", 0, ctypes.byref(f), ctypes.byref(b))
        ts.append((f.value, b.value))
    res[version] = c.host()
    print("version", version, "forward/backward ms", ts)
for f_ in FIELDS:
    print("   full size %-9s v2 vs v1 %.3e" % (f_, rel(res[2][f_], res[1][f_])))
# ---- timeline of v2 (block 0, backward steps 8..15)
_lib.call("khmm_debug_bw_backward_version", 2)
tr = torch.zeros(2 * 8 * 32, dtype=torch.int64, device="cuda")
L.khmm_debug_bw_trace(tr.data_ptr())
engine.estep_discrete_tc(p, ob)
torch.cuda.synchronize()
L.khmm_debug_bw_trace(None)
a = tr.cpu().numpy()[:256].reshape(8, 32)
names = {0: "tma:w_empty ok", 4: "mma:v_ready ok", 5: "mma:MMA1 issued", 6: "mma:wt_ready ok", 7: "mma:MMA2 issued",
         8: "epi:w_full ok", 9: "epi:reds issued", 10: "epi:d1_full ok", 11: "epi:chunk0 done", 12: "epi:v_ready sent",
         13: "epi:c0 e_full ok", 14: "epi:c0 loop done", 15: "epi:c0 mma2 group ok", 16: "red:slab0 ok", 17: "red:slab3 done"}
base = a[0, 8]
print("v2 step period cycles", (a[7, 12] - a[1, 12]) / 6)
for s_ in range(5, 8):
    ev = sorted((int(a[s_, k] - base), names[k]) for k in names if a[s_, k])
    print("step", s_ + 8, " | ".join("%s @%d" % (n, c) for c, n in ev))
