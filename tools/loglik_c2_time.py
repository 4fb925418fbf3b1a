"""Development: time of the fused small-N forward-backward log-likelihood kernel at config-2 shape, checked against float64.

    [KHMM_VARIANT=<name of a tools/build_variant.sh library>] [KHMM_SMALL_MMA=0] python tools/loglik_c2_time.py [sequences]"""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from kerehmm_b200 import _lib
if os.environ.get('KHMM_VARIANT'):
    _lib.LIB_PATH = os.path.join(_lib.PKG, 'libkhmm_%s.so' % os.environ['KHMM_VARIANT'])
import kerehmm_b200 as K
from kerehmm_b200 import engine
N, B, T = 8, int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 2048
np.random.seed(2)
h = K.GaussianHMM(N, 1, random_transitions=True, random_emissions=True)
mean = np.array([d.mean for d in h.emissionDistributions])
rng = np.random.default_rng(1002)
obs = [torch.from_numpy((mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T))).astype(np.float32)).cuda()
       for _ in range(2)]
dm = engine.DeviceModel(h._params(), torch.float32)
ll = torch.empty((B,), dtype=torch.float64, device='cuda'); llb = torch.empty_like(ll)
st = torch.cuda.current_stream()
def step(i):
    _lib.call("khmm_fwdbwd_loglik_gauss_f32", B, T, N, engine.ptr(dm.pi), engine.ptr(dm.A), engine.ptr(dm.mean),
              engine.ptr(dm.sd), engine.ptr(obs[i & 1]), _lib.F32, None, engine.ptr(ll), engine.ptr(llb), st.cuda_stream)
for i in range(5):
    step(i)
torch.cuda.synchronize()
ts = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        step(i)
    e1.record(); torch.cuda.synchronize()
    ts.append(round(e0.elapsed_time(e1) / 50, 4))
step(0); torch.cuda.synchronize()
ref = engine.forward(h._params(), engine.prepare_obs(obs[0][:256].double(), symbols=False), torch.float64, want_alpha=False,
                     want_scale=False)[2]
err = float(((ll[:256] - ref).abs() / ref.abs()).max())
print("variant", os.environ.get('KHMM_VARIANT'), "B", B, "ms/launch", ts, "fwd-bwd max rel diff", float(((ll - llb).abs() / ll.abs()).max()),
      "vs f64 (256 seq)", err, flush=True)
