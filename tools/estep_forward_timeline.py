"""Development: per-role clock64 timeline of the forward (and round-1 backward) tensor-core kernels (block 0, steps 8..15)."""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, '.')
import kerehmm_b200 as K
from kerehmm_b200 import engine, _lib
import os
if os.environ.get('KHMM_VARIANT'):
    _lib.LIB_PATH = os.path.join(_lib.PKG, 'libkhmm_%s.so' % os.environ['KHMM_VARIANT'])
N, M, B, T = 128, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 18944, 512
np.random.seed(4)
h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
obs = np.random.default_rng(1004).integers(0, M, size=(B, T)).astype(np.uint16)
ob = engine.prepare_obs(torch.from_numpy(obs.view(np.int16)).cuda(), symbols=True, u16_bits=True)
p = h._params()
engine.estep_discrete_tc(p, ob)
tr = torch.zeros(2 * 8 * 32, dtype=torch.int64, device="cuda")
L = _lib.lib()
L.khmm_debug_bw_trace(tr.data_ptr())
engine.estep_discrete_tc(p, ob)
torch.cuda.synchronize()
L.khmm_debug_bw_trace(None)
full = tr.cpu().numpy()
a = full[:256].reshape(8, 32)
f = full[256:].reshape(8, 32)
fn = {0: 'mma:a_ready ok', 1: 'mma:issued', 4: 'gat:buf_free ok', 5: 'gat:issued', 8: 'sto:w_ready ok', 9: 'sto:done', 12: 'epi:d_full ok', 13: 'epi:e_full ok', 14: 'epi:done'}
fb = f[0, 12]
for s_ in range(4, 8):
    ev = sorted((int(f[s_, k] - fb), fn[k]) for k in fn if f[s_, k])
    print('FWD step', s_ + 8, ' | '.join('%s @%d' % (n, c) for c, n in ev))
names = {0: "tma:alpha_empty ok", 4: "mma:v_ready ok", 5: "mma:MMA1 issued", 6: "mma:alpha_full ok", 7: "mma:MMA2 issued",
         8: "epi:d1_full ok", 9: "epi:alpha_full ok", 10: "epi:stage_empty ok", 11: "epi:passA done", 12: "epi:stage_full sent",
         13: "epi:mma2_done ok", 14: "epi:v_ready sent", 16: "red:stage_full ok", 17: "red:staging read"}
base = a[0, 8]
ev_t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev_t[0].record(); engine.estep_discrete_tc(p, ob); ev_t[1].record(); torch.cuda.synchronize()
print("variant", os.environ.get('KHMM_VARIANT'), "estep ms", ev_t[0].elapsed_time(ev_t[1]), "step period cycles", (a[7, 14] - a[1, 14]) / 6)
for s in range(6, 8):
    ev = sorted((int(a[s, k] - base), names[k]) for k in names if a[s, k])
    print("step", s + 8, " | ".join("%s @%d" % (n, c) for c, n in ev))
