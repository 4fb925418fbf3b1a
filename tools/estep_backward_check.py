"""Development: the quarter-tile pipeline backward kernel (v3, the default) against the round-1 kernel (v1) and the float64
oracle on small shapes, both operand precisions; then kernel times at config-4 size and a clock64 timeline of block 0.

    python tools/estep_backward_check.py [sequences, default 131072]      (QUICK=1: one small shape only)"""
import ctypes, sys, os, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine, _lib
from oracle import hmm_oracle as O

L = _lib.lib()
FIELDS = ("pi", "gam", "emis_den", "emis_num", "xi_raw", "loglik", "nseq", "flags")
PRECS = {"tf32": engine.TC_TF32, "bf16x3": engine.TC_BF16X3}


def run(version, p, ob, prec):
    _lib.call("khmm_debug_bw_backward_version", version)
    c = engine.estep_discrete_tc(p, ob, precision=prec).host()
    torch.cuda.synchronize()
    return c


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-12 * max(1.0, np.abs(b).max()))
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


SHAPES = [(128, 16, 128, 2, False), (128, 16, 128, 3, False), (128, 16, 130, 8, False), (128, 16, 64, 1, False),
          (128, 48, 300, 33, True), (128, 1024, 1000, 20, False), (128, 64, 148 * 128 * 2 + 77, 6, True)]
if os.environ.get("QUICK"):
    SHAPES = SHAPES[4:5]
ok = True
for (N, M, B, T, ragged) in SHAPES:
    np.random.seed(N + M + B + T)
    h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
    pi, A, Bm = h.initialProbabilities.copy(), h.transitionMatrix.copy(), h.emission_matrix()
    rng = np.random.default_rng(B * T)
    obs = rng.integers(0, M, size=(B, T))
    lengths = rng.integers(1, T + 1, size=B) if ragged else None
    ob = engine.prepare_obs(obs, symbols=True, lengths=lengths, alphabet=M)
    p = h._params()
    st = None
    if B <= 2000:
        if lengths is None:
            st = O.batch_stats_discrete(pi, A, Bm, obs)
        else:
            for b in range(B):
                s1 = O.batch_stats_discrete(pi, A, Bm, obs[b:b + 1, :lengths[b]])
                st = s1 if st is None else {k: st[k] + s1[k] for k in st}
    print("shape N=%d M=%d B=%d T=%d ragged=%s" % (N, M, B, T, ragged), flush=True)
    for pname, prec in PRECS.items():
        try:
            c1 = run(1, p, ob, prec)
            c3 = run(3, p, ob, prec)
        except Exception as e:
            print("   ", pname, "FAILED:", e)
            ok = False
            break
        ref = None
        if st is not None:
            ref = dict(pi=st["pi"], gam=st["gam"], emis_den=st["emis_den"], emis_num=st["emis_num"],
                       xi_raw=st["xi"] / np.where(A > 0, A, 1.0), loglik=st["loglik"], nseq=st["nseq"], flags=0.0)
        for f in FIELDS:
            print("   %-6s %-9s v3 vs v1 %.3e   %s   finite %s" % (
                pname, f, rel(c3[f], c1[f]),
                ("v1 vs oracle %.3e   v3 vs oracle %.3e" % (rel(c1[f], ref[f]), rel(c3[f], ref[f]))) if ref else "",
                bool(np.isfinite(np.asarray(c3[f])).all())), flush=True)
    if not ok:
        break

# ---- timing at config-4 size
N, M, B, T = 128, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 131072, 512
np.random.seed(4)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
obs = np.random.default_rng(1004).integers(0, M, size=(B, T)).astype(np.uint16)
ob = engine.prepare_obs(torch.from_numpy(obs.view(np.int16)).cuda(), symbols=True, u16_bits=True)
p = h._params()
_lib.call("khmm_estep_tc_timing", 1)
res = {}
for pname, prec in PRECS.items():
    for version in (1, 3):
        _lib.call("khmm_debug_bw_backward_version", version)
        for _ in range(2):
            c = engine.estep_discrete_tc(p, ob, precision=prec)
        torch.cuda.synchronize()
        f, b = ctypes.c_float(), ctypes.c_float()
        ts = []
        for _ in range(3):
            c = engine.estep_discrete_tc(p, ob, precision=prec)
            _lib.call("khmm_estep_tc_kernel_ms", 0, ctypes.byref(f), ctypes.byref(b))
            ts.append((round(f.value, 2), round(b.value, 2)))
        res[(pname, version)] = c.host()
        print(pname, "version", version, "forward/backward ms", ts, flush=True)
    for f_ in FIELDS:
        print("   full size %-6s %-9s v3 vs v1 %.3e" % (pname, f_, rel(res[(pname, 3)][f_], res[(pname, 1)][f_])))
# ---- timeline of v3 (block 0, backward steps 8..15)
names = {0: "tma:w_empty0 ok", 4: "mma:v_ready0 ok", 5: "mma:MMA1.0 issued", 6: "mma:w_full0 ok", 7: "mma:MMA2.0 issued",
         20: "mma:v_ready1 ok", 21: "mma:MMA1.1 issued", 22: "mma:w_full1 ok", 23: "mma:MMA2.1 issued",
         8: "c0:prep ok", 10: "c0:d1_full ok", 9: "c0:w_full ok", 11: "c0:next prep ok", 13: "c0:ldtm done", 14: "c0:body done",
         15: "c0:mma2_done ok", 12: "c0:v_ready sent", 24: "c1:d1_full ok", 26: "c1:v_ready sent"}
for pname, prec in PRECS.items():
    _lib.call("khmm_debug_bw_backward_version", 3)
    tr = torch.zeros(2 * 8 * 32, dtype=torch.int64, device="cuda")
    L.khmm_debug_bw_trace(tr.data_ptr())
    engine.estep_discrete_tc(p, ob, precision=prec)
    torch.cuda.synchronize()
    L.khmm_debug_bw_trace(None)
    a = tr.cpu().numpy()[:256].reshape(8, 32)
    base = a[0, 8]
    print(pname, "v3 step period cycles", (a[7, 12] - a[1, 12]) / 6)
    for s_ in range(5, 8):
        ev = sorted((int(a[s_, k] - base), names[k]) for k in names if a[s_, k])
        print("step", s_ + 8, " | ".join("%s @%d" % (n, c) for c, n in ev))
_lib.call("khmm_debug_bw_backward_version", 0)
