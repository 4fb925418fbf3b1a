"""Device-side glue between the Python HMM classes and libkhmm (include/khmm.h).

PyTorch is used here only for device memory, streams and host<->device copies; every
computation on the hot path is a libkhmm kernel.  There is no CPU fallback: without a CUDA
device or without the built library these functions raise.

Input rule for the batched API: observations given as a NumPy array / CPU tensor ("host
buffers") are copied to the device inside the call and results come back as NumPy arrays;
observations given as a CUDA tensor are used in place and results stay on the device as torch
tensors.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise RuntimeError("kerehmm_b200 needs a CUDA device (built for B200 / sm_100a); "
                           "there is no CPU fallback.")
    _lib.lib()
    return t


def device():
    t = require_cuda()
    return t.device("cuda", t.cuda.current_device())


def stream_ptr():
    return torch().cuda.current_stream().cuda_stream


def ptr(x):
    return None if x is None else x.data_ptr()


# ------------------------------------------------------------------------- model packing
@dataclass
class ModelParams:
    """Host float64 parameters of one HMM, in the reference's conventions."""
    kind: str                       # "discrete" | "gauss" | "gauss_full"
    pi: np.ndarray                  # (N,)
    A: np.ndarray                   # (N, N)
    B: Optional[np.ndarray] = None  # discrete: (N, M) state-major, as the reference stores it
    mean: Optional[np.ndarray] = None   # gauss: (N,) ; gauss_full: (N, D)
    sd: Optional[np.ndarray] = None     # gauss: (N,) -- the reference's `variance` attribute
    cov: Optional[np.ndarray] = None    # gauss_full: (N, D, D)

    @property
    def N(self):
        return int(self.A.shape[0])

    @property
    def M(self):
        return int(self.B.shape[1])


def _dev(a, dtype):
    t = torch()
    return t.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device(), non_blocking=False)


class DeviceModel:
    """Device copies of the parameters in one real type (re-made on every call: the reference's
    parameters are plain mutable arrays, SURVEY.md 8(b) ownership)."""

    def __init__(self, p: ModelParams, real):
        t = torch()
        self.p, self.real = p, real
        self.suffix = "f32" if real == t.float32 else "f64"
        self.pi = _dev(p.pi, real)
        self.A = _dev(p.A, real)
        # transposes are taken on the device: a strided host copy of a 1024 x 1024 float64 matrix plus a second
        # upload cost ~3 ms per call at config 5
        self.AT = self.A.t().contiguous()
        if p.kind == "discrete":
            self.Bsm = _dev(p.B, real).t().contiguous()          # symbol-major [M][N]
        elif p.kind == "gauss":
            self.mean = _dev(p.mean, real)
            self.sd = _dev(p.sd, real)


_INT_CODES = {"uint8": _lib.U8, "int16": _lib.U16, "uint16": _lib.U16, "int32": _lib.I32, "int64": _lib.I64}


@dataclass
class Obs:
    tensor: object      # torch tensor on the device
    code: int           # libkhmm dtype code
    B: int
    T: int
    on_host: bool       # the caller passed host memory -> results go back to the host
    lengths: object = None
    h2d_bytes: int = 0
    # overlapped upload (prepare_obs(..., overlap=True)): row blocks still in flight on a side stream --
    # [(first row, end row, event, page-locked (min, max) of the block's symbols or None)] -- and the alphabet they are
    # checked against
    ready: object = None
    alphabet: int = None
    starter: object = None      # queues the copies; called by the first wait_block, i.e. AFTER the consumer's own (small)
                                # parameter uploads -- behind the batch they would wait for all of it on the copy engine

    def wait_block(self, k):
        """Make the current stream wait for row block k of an overlapped upload and run its deferred symbol check
        (raises the reference's IndexError).  -> (first row, end row)."""
        t = torch()
        if self.starter is not None:
            self.starter(self)
            self.starter = None
        b0, b1, ev, cell = self.ready[k]
        t.cuda.current_stream().wait_event(ev)
        if cell is not None:
            ev.synchronize()                        # the copy, the range kernel and its read-back of THIS block only
            lo, hi = (int(v) for v in cell.numpy())
            if lo <= hi and (hi >= self.alphabet or lo < 0):
                raise IndexError("index %d is out of bounds for axis 0 with size %d" % (hi if hi >= self.alphabet else lo,
                                                                                       self.alphabet))
        return b0, b1

    def wait_all(self):
        """Consumers that do not work block by block: wait for the whole upload."""
        if self.ready is not None:
            for k in range(len(self.ready)):
                self.wait_block(k)
            self.ready = None


def _as_tensor(x):
    t = torch()
    if isinstance(x, t.Tensor):
        return x
    a = np.asarray(x)
    if a.dtype == np.uint16:            # torch has no first-class uint16: ship the bits as int16
        a = a.view(np.int16)
    if a.dtype == np.int8:
        a = a.astype(np.int16)
    if a.dtype.kind == "u" and a.dtype.itemsize > 2:
        a = a.astype(np.int64)
    if a.dtype == np.bool_ or a.dtype.kind not in "iuf":
        raise TypeError("observations must be numeric, got %s" % a.dtype)
    return t.from_numpy(np.ascontiguousarray(a))


_UNSIGNED_LIMIT = {_lib.U8: 256, _lib.U16: 65536}


def check_symbols(x, code: int, B: int, T: int, lengths, alphabet: int):
    """The reference indexes a NumPy table with every symbol (distribution.py:35-36): a symbol >= M or < -M raises
    IndexError and a negative one wraps to M + s.  The kernels only clamp (so that padding can hold anything), hence
    the live positions are range-checked here, once per batch, with one device reduction (`khmm_symbol_range`).
    -> the tensor to use (a wrapped copy when negative symbols are present)."""
    if B == 0 or (code in _UNSIGNED_LIMIT and alphabet >= _UNSIGNED_LIMIT[code]):
        return x                      # nothing to check / every representable value is a valid symbol
    t = torch()
    mm = t.empty((2,), dtype=t.int64, device=x.device)
    _lib.call("khmm_symbol_range", B, T, ptr(x), code, ptr(lengths), ptr(mm), stream_ptr())
    lo, hi = (int(v) for v in mm.cpu().numpy())
    if B == 0 or lo > hi:
        return x
    if hi >= alphabet or lo < -alphabet:
        bad = hi if hi >= alphabet else lo
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (bad, alphabet))
    if lo < 0:                        # NumPy wraps negative indices
        x = t.where(x < 0, x + alphabet, x)
    return x


OVERLAP_MIN_BYTES = 32 << 20      # below this an overlapped upload is not worth its extra launches
_side_streams = {}                # device index -> the upload stream of the overlapped path


def training_blocks(B: int):
    """Row blocks of an overlapped upload for the tensor-core E-step: one tile of 128 sequences per SM, two per SM, the
    rest -- so that the block-wise E-step runs the same number of tile rounds as one call over the whole batch."""
    r = 128 * torch().cuda.get_device_properties(device()).multi_processor_count
    return [r, 3 * r, B] if B > 4 * r else None


def decoding_blocks(B: int):
    """Row blocks for Viterbi (a warp per sequence, eight warps per CTA; r = two CTA rounds): a small first block so that
    decoding starts early and a small last one so that only its paths are still to be copied back when the kernels end,
    each with a three times larger neighbour -- with eight ranks copying at once the bulk of the batch takes longer to
    arrive (and its paths longer to leave) than the small block alone takes to decode."""
    r = 2 * 8 * torch().cuda.get_device_properties(device()).multi_processor_count
    return [r, 4 * r, B - 4 * r, B - r, B] if B > 16 * r else None


def _upload_overlapped(x, code: int, alphabet, ends):
    """Page-locked host batch of unsigned symbols -> Obs whose row blocks (ending at `ends`) arrive on a side stream while
    the consumer already works on the first ones."""
    t = torch()
    dev = device()
    B, T = int(x.shape[0]), int(x.shape[1])
    bounds = [0] + list(ends)
    nblk = len(ends)
    # ONE side stream per device, kept: a new stream per call has an empty allocator pool, and the cudaMalloc that
    # fills it (1 - 7 ms, measured) would sit in front of the first kernel launch.  High priority: its small range
    # kernels go in front of queued E-step CTAs.
    side = _side_streams.get(dev.index)
    if side is None:
        side = _side_streams[dev.index] = t.cuda.Stream(device=dev, priority=-1)
    xd = t.empty((B, T), dtype=x.dtype, device=dev)
    check = alphabet is not None and not (code in _UNSIGNED_LIMIT and alphabet >= _UNSIGNED_LIMIT[code])
    mm = t.empty((nblk, 2), dtype=t.int64, device=dev) if check else None       # allocated from the main stream's pool
    cells = t.empty((nblk, 2), dtype=t.int64, pin_memory=True) if check else None
    ob = Obs(tensor=xd, code=code, B=B, T=T, on_host=True, lengths=None, h2d_bytes=x.numel() * x.element_size(),
             ready=[(bounds[k], bounds[k + 1], None, None) for k in range(nblk)],
             alphabet=None if alphabet is None else int(alphabet))

    def start(ob):
        side.wait_stream(t.cuda.current_stream())
        ready = []
        with t.cuda.stream(side):
            for k in range(nblk):
                b0, b1 = bounds[k], bounds[k + 1]
                xd[b0:b1].copy_(x[b0:b1], non_blocking=True)
                cell = None
                if check:
                    _lib.call("khmm_symbol_range", b1 - b0, T, ptr(xd[b0:b1]), code, None, ptr(mm[k]), side.cuda_stream)
                    cell = cells[k]
                    cell.copy_(mm[k], non_blocking=True)   # read back on the side stream, before the block's event
                ev = t.cuda.Event()
                ev.record(side)
                ready.append((b0, b1, ev, cell))
        xd.record_stream(side)
        if check:
            mm.record_stream(side)
        ob.ready = ready

    ob.starter = start
    return ob


def prepare_obs(obs, symbols: bool, lengths=None, feature_dims: int = 0, u16_bits: bool = False,
                alphabet: int = None, overlap=None) -> Obs:
    """Move a batch of observations to the device.  symbols=True: integer symbols (B, T);
    else real observations (B, T) or (B, T, D).  `alphabet` (discrete models): the table size M the live symbols are
    checked against, with the reference's IndexError / negative-index behaviour.  `overlap` (`training_blocks` /
    `decoding_blocks`): the caller consumes the batch block by block (`Obs.ready`): a large page-locked batch of unsigned
    symbols is then copied on a side stream while the first blocks are already being worked on, its symbol check
    deferred to each block."""
    t = require_cuda()
    was_u16 = (not isinstance(obs, t.Tensor)) and np.asarray(obs).dtype == np.uint16
    x = _as_tensor(obs)
    on_host = not x.is_cuda
    want_nd = 3 if feature_dims else 2
    if x.dim() != want_nd:
        raise ValueError("expected observations with %d dims (batch, time%s), got shape %s"
                         % (want_nd, ", features" if feature_dims else "", tuple(x.shape)))
    if symbols:
        if x.dtype.is_floating_point:
            # the reference indexes tables with the observation, which fails for floats (Q10)
            raise IndexError("only integers are valid symbols for a discrete emission table")
        name = str(x.dtype).replace("torch.", "")
        if name not in _INT_CODES:
            x = x.to(t.int64)
            name = "int64"
        code = _INT_CODES[name]
        if name == "int16" and not (was_u16 or u16_bits):
            x = x.to(t.int32)
            code = _lib.I32
    else:
        if not x.dtype.is_floating_point:
            x = x.to(t.float32)
        if x.dtype not in (t.float32, t.float64):
            x = x.to(t.float32)
        code = _lib.F32 if x.dtype == t.float32 else _lib.F64
    h2d = 0
    if (overlap is not None and on_host and symbols and lengths is None and code in _UNSIGNED_LIMIT and x.dim() == 2
            and x.is_contiguous() and x.numel() * x.element_size() >= OVERLAP_MIN_BYTES and x.is_pinned()
            and int(x.shape[1]) >= 1):
        ends = overlap(int(x.shape[0]))
        if ends is not None:
            return _upload_overlapped(x, code, alphabet, ends)
    if on_host:
        h2d = x.numel() * x.element_size()
        x = x.to(device(), non_blocking=True)
    x = x.contiguous()
    B, T = int(x.shape[0]), int(x.shape[1])
    if T < 1:
        raise ValueError("sequences must have at least one observation")
    L = None
    if lengths is not None:
        la = np.asarray(lengths.cpu() if isinstance(lengths, t.Tensor) else lengths).astype(np.int64)
        if la.shape != (B,):
            raise ValueError("lengths must have shape (B,)")
        if B and (la.min() < 1 or la.max() > T):
            raise ValueError("lengths must lie in [1, T]")
        L = t.as_tensor(la.astype(np.int32)).to(device())
    if symbols and alphabet is not None:
        x = check_symbols(x, code, B, T, L, int(alphabet))
    return Obs(tensor=x, code=code, B=B, T=T, on_host=on_host, lengths=L, h2d_bytes=h2d)


_PINNED_RESULT_MIN_BYTES = 1 << 20


def _to_host(x):
    """Device tensor -> NumPy.  Large results (a config-3 path array is 268 MB) land in page-locked memory from
    torch's caching host allocator and the returned array keeps that block alive: a pageable `.cpu()` copy of that
    size costs ~100 ms of page faults and staging, the pinned copy ~5 ms at PCIe rate; the block goes back to the
    allocator's cache when the caller drops the array."""
    t = torch()
    if x.numel() * x.element_size() < _PINNED_RESULT_MIN_BYTES:
        return x.cpu().numpy()
    try:
        host = t.empty(x.shape, dtype=x.dtype, pin_memory=True)
    except RuntimeError:          # page-locked memory exhausted: fall back to a pageable copy
        return x.cpu().numpy()
    host.copy_(x, non_blocking=True)
    t.cuda.current_stream().synchronize()
    return host.numpy()


def finish(ob: Obs, *tensors):
    """Return results on the side the observations came from."""
    if ob.on_host:
        out = tuple(None if x is None else (x.array if isinstance(x, HostResult) else _to_host(x)) for x in tensors)
    else:
        out = tensors
    return out if len(out) != 1 else out[0]


# ------------------------------------------------------------------------------ kernels
def _emission_args(dm: DeviceModel):
    k = dm.p.kind
    if k == "discrete":
        return "discrete", (dm.p.M,), (ptr(dm.Bsm),)
    if k == "gauss":
        return "gauss", (), (ptr(dm.mean), ptr(dm.sd))
    raise ValueError(k)


GAUSS_FULL_DMAX = 8     # KHMM_GAUSS_FULL_DMAX: feature count up to which full-covariance emissions are fused into the recursions


def _gauss_full_operands(p: ModelParams):
    """Device float64 (means[N][D], Linv[N][D][D], lognorm[N]) of a full-covariance model: the inverse lower Cholesky
    factor and the log normaliser of each state's density (distribution.py:153)."""
    t = torch()
    D = p.mean.shape[1]
    Ls = np.linalg.cholesky(p.cov)
    Linv = np.stack([np.linalg.inv(L) for L in Ls])
    lognorm = -0.5 * (D * np.log(2.0 * np.pi) + 2.0 * np.log(np.diagonal(Ls, axis1=1, axis2=2)).sum(axis=1))
    return D, _dev(p.mean, t.float64), _dev(Linv, t.float64), _dev(lognorm, t.float64)


def dense_emissions(p: ModelParams, ob: Obs):
    """E[B][T][N] float64 on the device (abstract_hmm.py:362-366 batched)."""
    t = require_cuda()
    N = p.N
    E = t.empty((ob.B, ob.T, N), dtype=t.float64, device=device())
    if p.kind == "discrete":
        Bsm = _dev(p.B.T, t.float64)
        _lib.call("khmm_emissions_discrete_f64", ob.B, ob.T, N, p.M, ptr(Bsm), ptr(ob.tensor), ob.code, ptr(E),
                  stream_ptr())
    elif p.kind == "gauss":
        mean, sd = _dev(p.mean, t.float64), _dev(p.sd, t.float64)
        _lib.call("khmm_emissions_gauss_f64", ob.B, ob.T, N, ptr(mean), ptr(sd), ptr(ob.tensor), ob.code, ptr(E),
                  stream_ptr())
    else:
        D, dm_, dl, dn = _gauss_full_operands(p)
        _lib.call("khmm_emissions_gauss_full_f64", ob.B, ob.T, N, D, ptr(dm_), ptr(dl), ptr(dn), ptr(ob.tensor),
                  ob.code, ptr(E), stream_ptr())
    return E


def forward(p: ModelParams, ob: Obs, real, want_alpha=True, want_scale=True, want_loglik=True):
    """Scaled forward on the device -> (alpha|None, scale|None, loglik|None) device tensors."""
    t = require_cuda()
    dm = DeviceModel(p, real)
    N, dev = p.N, device()
    alpha = t.empty((ob.B, ob.T, N), dtype=real, device=dev) if want_alpha else None
    scale = t.empty((ob.B, ob.T), dtype=real, device=dev) if want_scale else None
    loglik = t.empty((ob.B,), dtype=t.float64, device=dev) if want_loglik else None
    if p.kind == "gauss_full" and p.mean.shape[1] <= GAUSS_FULL_DMAX:
        # fused emissions (round 2): no [B][T][N] emission matrix
        D, dmean, dlinv, dnorm = _gauss_full_operands(p)
        _lib.call("khmm_forward_gaussfull_" + dm.suffix, ob.B, ob.T, N, D, ptr(dm.pi), ptr(dm.A), ptr(dmean), ptr(dlinv),
                  ptr(dnorm), ptr(ob.tensor), ob.code, ptr(ob.lengths), ptr(alpha), ptr(scale), ptr(loglik), stream_ptr())
        return alpha, scale, loglik
    if p.kind == "gauss_full":
        E = dense_emissions(p, ob).to(real)
        _lib.call("khmm_forward_dense_" + dm.suffix, ob.B, ob.T, N, ptr(dm.pi), ptr(dm.A), ptr(E), ptr(ob.lengths),
                  ptr(alpha), ptr(scale), ptr(loglik), stream_ptr())
        return alpha, scale, loglik
    name, ints, ptrs = _emission_args(dm)
    _lib.call("khmm_forward_%s_%s" % (name, dm.suffix), ob.B, ob.T, N, *ints, ptr(dm.pi), ptr(dm.A), *ptrs,
              ptr(ob.tensor), ob.code, ptr(ob.lengths), ptr(alpha), ptr(scale), ptr(loglik), stream_ptr())
    return alpha, scale, loglik


def backward(p: ModelParams, ob: Obs, real, scale, alpha=None, want_beta=True, want_gamma=False,
             gamma_kind=_lib.GAMMA_REFERENCE):
    """Scaled backward (+ optional fused gamma) -> (beta|None, gamma|None)."""
    t = require_cuda()
    if tuple(scale.shape) != (ob.B, ob.T):
        raise ValueError("scale must have shape (B, T) = %s, got %s" % ((ob.B, ob.T), tuple(scale.shape)))
    if alpha is not None and tuple(alpha.shape) != (ob.B, ob.T, p.N):
        raise ValueError("alpha must have shape (B, T, N)")
    dm = DeviceModel(p, real)
    N, dev = p.N, device()
    beta = t.empty((ob.B, ob.T, N), dtype=real, device=dev) if want_beta else None
    gamma = t.empty((ob.B, ob.T, N), dtype=real, device=dev) if want_gamma else None
    if p.kind == "gauss_full" and p.mean.shape[1] <= GAUSS_FULL_DMAX:
        D, dmean, dlinv, dnorm = _gauss_full_operands(p)
        _lib.call("khmm_backward_gaussfull_" + dm.suffix, ob.B, ob.T, N, D, ptr(dm.AT), ptr(dmean), ptr(dlinv), ptr(dnorm),
                  ptr(ob.tensor), ob.code, ptr(ob.lengths), ptr(scale), ptr(alpha), ptr(beta), ptr(gamma), gamma_kind,
                  stream_ptr())
        return beta, gamma
    if p.kind == "gauss_full":
        E = dense_emissions(p, ob).to(real)
        _lib.call("khmm_backward_dense_" + dm.suffix, ob.B, ob.T, N, ptr(dm.AT), ptr(E), ptr(ob.lengths), ptr(scale),
                  ptr(alpha), ptr(beta), ptr(gamma), gamma_kind, stream_ptr())
        return beta, gamma
    name, ints, ptrs = _emission_args(dm)
    _lib.call("khmm_backward_%s_%s" % (name, dm.suffix), ob.B, ob.T, N, *ints, ptr(dm.AT), *ptrs, ptr(ob.tensor),
              ob.code, ptr(ob.lengths), ptr(scale), ptr(alpha), ptr(beta), ptr(gamma), gamma_kind, stream_ptr())
    return beta, gamma


def fwdbwd_loglik_small(p: ModelParams, ob: Obs, both=True):
    """Fused forward(+backward) log-likelihood for N <= 16, fp32 -> (loglik, loglik_bwd|None)."""
    t = require_cuda()
    if p.N > _lib.SMALL_N_MAX:
        raise ValueError("fused log-likelihood kernel covers N <= %d" % _lib.SMALL_N_MAX)
    dm = DeviceModel(p, t.float32)
    dev = device()
    ll = t.empty((ob.B,), dtype=t.float64, device=dev)
    llb = t.empty((ob.B,), dtype=t.float64, device=dev) if both else None
    if p.kind == "gauss":
        _lib.call("khmm_fwdbwd_loglik_gauss_f32", ob.B, ob.T, p.N, ptr(dm.pi), ptr(dm.A), ptr(dm.mean), ptr(dm.sd),
                  ptr(ob.tensor), ob.code, ptr(ob.lengths), ptr(ll), ptr(llb), stream_ptr())
    elif p.kind == "discrete":
        _lib.call("khmm_fwdbwd_loglik_discrete_f32", ob.B, ob.T, p.N, p.M, ptr(dm.pi), ptr(dm.A), ptr(dm.Bsm),
                  ptr(ob.tensor), ob.code, ptr(ob.lengths), ptr(ll), ptr(llb), stream_ptr())
    else:
        raise ValueError("fused log-likelihood kernel needs discrete or 1-D Gaussian emissions")
    return ll, llb


def fwdbwd_loglik_small_host(p: ModelParams, obs, both=True, nchunks=4, min_bytes=16 << 20):
    """Host-buffer variant of `fwdbwd_loglik_small`: the batch is cut into `nchunks` row blocks whose
    host->device copies run on a side stream while the kernel works on the previous block, so the call
    costs about max(copy, compute) instead of their sum.  Returns None when the input is not a large,
    equal-length host batch (the caller then takes the plain path).  -> (loglik, loglik_bwd|None) NumPy."""
    t = require_cuda()
    if isinstance(obs, t.Tensor):
        x = obs
    else:
        a = np.asarray(obs)
        if a.dtype == np.uint16 or a.dtype.kind not in "iuf" or a.dtype == np.int8:
            return None
        x = t.from_numpy(np.ascontiguousarray(a))
    symbols = p.kind == "discrete"
    if x.is_cuda or x.dim() != 2 or x.numel() * x.element_size() < min_bytes or p.N > _lib.SMALL_N_MAX:
        return None
    if symbols:
        name = str(x.dtype).replace("torch.", "")
        if name not in ("uint8", "int32", "int64"):
            return None
        code = _INT_CODES[name]
    else:
        if x.dtype not in (t.float32, t.float64):
            return None
        code = _lib.F32 if x.dtype == t.float32 else _lib.F64
    x = x.contiguous()
    B, T = int(x.shape[0]), int(x.shape[1])
    dev = device()
    main = t.cuda.current_stream()
    side = t.cuda.Stream(device=dev)
    xd = t.empty((B, T), dtype=x.dtype, device=dev)
    res = t.empty((2 if both else 1, B), dtype=t.float64, device=dev)      # one device->host copy for both results
    ll, llb = res[0], (res[1] if both else None)
    bounds = [(B * k) // nchunks for k in range(nchunks + 1)]
    side.wait_stream(main)
    events = []
    with t.cuda.stream(side):
        for k in range(nchunks):
            b0, b1 = bounds[k], bounds[k + 1]
            if b1 > b0:
                xd[b0:b1].copy_(x[b0:b1], non_blocking=True)
            ev = t.cuda.Event()
            ev.record(side)
            events.append(ev)
    # the (small, blocking) parameter uploads go behind the first block's copy instead of in front of it
    dm = DeviceModel(p, t.float32)
    for k in range(nchunks):
        b0, b1 = bounds[k], bounds[k + 1]
        if b1 <= b0:
            continue
        main.wait_event(events[k])
        lk, lbk = ll[b0:b1], (llb[b0:b1] if both else None)
        if p.kind == "gauss":
            _lib.call("khmm_fwdbwd_loglik_gauss_f32", b1 - b0, T, p.N, ptr(dm.pi), ptr(dm.A), ptr(dm.mean), ptr(dm.sd),
                      ptr(xd[b0:b1]), code, None, ptr(lk), ptr(lbk), main.cuda_stream)
        else:
            _lib.call("khmm_fwdbwd_loglik_discrete_f32", b1 - b0, T, p.N, p.M, ptr(dm.pi), ptr(dm.A), ptr(dm.Bsm),
                      ptr(xd[b0:b1]), code, None, ptr(lk), ptr(lbk), main.cuda_stream)
    xd.record_stream(main)
    host = res.cpu().numpy()
    return host[0], (host[1] if both else None)


TC_MIN_STATES = 65      # below this the padding to 128 states costs more than the float32 SIMT kernels


def tc_supported(p: ModelParams, ob: Obs) -> bool:
    """The tcgen05 recursion covers 65 <= N <= 1024 (state spaces that are not a multiple of 128 are padded with
    unreachable states), equal-length or ragged batches."""
    return p.kind in ("discrete", "gauss") and TC_MIN_STATES <= p.N <= 1024


def pad_states(p: ModelParams, multiple: int = 128) -> ModelParams:
    """The same model with its state space padded to a multiple of `multiple` by UNREACHABLE states: start
    probability 0, no transitions in or out, and emission 0 (a zero table row; a Gaussian with sd = inf, whose density
    the kernels evaluate as exp2(-inf) = 0).  Their alpha and beta are exactly 0 at every step, every likelihood and
    every scale factor is unchanged."""
    N = p.N
    Np = -(-N // multiple) * multiple
    if Np == N:
        return p
    pi = np.zeros(Np); pi[:N] = p.pi
    A = np.zeros((Np, Np)); A[:N, :N] = p.A
    if p.kind == "discrete":
        B = np.zeros((Np, p.M)); B[:N] = p.B
        return ModelParams(kind="discrete", pi=pi, A=A, B=B)
    mean = np.zeros(Np); mean[:N] = p.mean
    sd = np.full(Np, np.inf); sd[:N] = p.sd
    return ModelParams(kind="gauss", pi=pi, A=A, mean=mean, sd=sd)


TC_TF32, TC_BF16, TC_BF16X3 = 0, 1, 2

# Operand precision of the fused tensor-core E-step when the caller just says tensor_cores=True / "auto":
# split bf16 operands (three MMAs, 2^-17) -- counts and re-estimated parameters then agree with the float64
# reference to float32 accuracy; "tf32" (one MMA, 2^-11, ~8 % faster per iteration) stays available by name.
ESTEP_TC_DEFAULT_PRECISION = TC_BF16X3


def estep_precision(tensor_cores):
    """tensor_cores argument of the Baum-Welch entry points -> (use the tensor-core path?, operand precision).
    False | True | "auto" | "tf32" | "bf16x3"."""
    if tensor_cores in ("tf32",):
        return True, TC_TF32
    if tensor_cores in ("bf16x3", "split"):
        return True, TC_BF16X3
    if tensor_cores in (True, False, "auto"):
        return tensor_cores, ESTEP_TC_DEFAULT_PRECISION
    raise ValueError('tensor_cores must be True, False, "auto", "tf32" or "bf16x3"')


def fwdbwd_loglik_tc(p: ModelParams, ob: Obs, both=True, precision=TC_TF32):
    """Tensor-core (tcgen05) forward(+backward) log-likelihood -> (loglik, loglik_bwd|None).
    precision: TC_TF32 (default) or TC_BF16 (bf16 operands, fp32 accumulation; twice the MMA rate; tf32 is used
    instead on a device without cooperative launch, the one case the bf16 kernel does not cover)."""
    t = require_cuda()
    p = pad_states(p)               # N not a multiple of 128: unreachable padding states
    dm = DeviceModel(p, t.float32)
    dev = device()
    ll = t.empty((ob.B,), dtype=t.float64, device=dev)
    llb = t.empty((ob.B,), dtype=t.float64, device=dev) if both else None
    nbytes = _lib.lib().khmm_fwdbwd_loglik_tc_workspace_bytes(ob.B, p.N)
    ws, wsp = workspace(nbytes, 256)
    def run(prec):
        if p.kind == "gauss":
            _lib.call("khmm_fwdbwd_loglik_tc_gauss_ragged_f32", ob.B, ob.T, p.N, ptr(dm.pi), ptr(dm.A), ptr(dm.AT),
                      ptr(dm.mean), ptr(dm.sd), ptr(ob.tensor), ob.code, ptr(ob.lengths), wsp, nbytes, ptr(ll), ptr(llb),
                      prec, stream_ptr())
        else:
            _lib.call("khmm_fwdbwd_loglik_tc_discrete_ragged_f32", ob.B, ob.T, p.N, p.M, ptr(dm.pi), ptr(dm.A), ptr(dm.AT),
                      ptr(dm.Bsm), ptr(ob.tensor), ob.code, ptr(ob.lengths), wsp, nbytes, ptr(ll), ptr(llb), prec,
                      stream_ptr())
    try:
        run(precision)
    except _lib.KhmmError as e:
        if precision == TC_BF16 and e.code == _lib.ERR_UNSUPPORTED:
            import warnings
            warnings.warn("fwdbwd_loglik_tc: bf16 operands are not available for this shape / device (%s); running with tf32 "
                          "operands (more precise, about half the tensor rate)" % e, RuntimeWarning, stacklevel=2)
            run(TC_TF32)
        else:
            raise
    return ll, llb


_PATH_CODES = {"uint8": _lib.U8, "int16": _lib.U16, "int32": _lib.I32, "int64": _lib.I64}


def viterbi(p: ModelParams, ob: Obs, path_dtype=None):
    """Bit-exact float64 Viterbi -> (path[B][T], logp[B]) device tensors.  The log tables are
    taken on the HOST with np.log, exactly what abstract_hmm.py:218,227,229 evaluates."""
    t = require_cuda()
    if p.kind not in ("discrete", "gauss"):
        raise IndexError("viterbi needs integer observations (discrete emissions) or 1-D Gaussian emissions "
                         "(batched extension); the reference fails on float observations (abstract_hmm.py:234-237)")
    N, dev = p.N, device()
    with np.errstate(divide="ignore"):
        logpi, logA = np.log(p.pi), np.log(p.A)
    d_logpi, d_logA = _dev(logpi, t.float64), _dev(logA, t.float64)
    path_dtype = path_dtype or t.int64
    pname = str(path_dtype).replace("torch.", "")
    if pname not in _PATH_CODES:
        raise ValueError("path_dtype must be uint8, int16, int32 or int64")
    capacity = {"uint8": 256, "int16": 32768}.get(pname)
    if capacity is not None and N > capacity:
        raise ValueError("path_dtype %s cannot hold the states of a %d-state model" % (pname, N))
    path = t.zeros((ob.B, ob.T), dtype=path_dtype, device=dev)
    logp = t.empty((ob.B,), dtype=t.float64, device=dev)
    # an overlapped upload (prepare_obs(..., overlap=decoding_blocks)) is decoded block by block, and the paths of a
    # block travel back on a second side stream while the next block is decoded
    blocks = [(0, ob.B)] if ob.ready is None else [None] * len(ob.ready)
    bmax = ob.B if ob.ready is None else max(b1 - b0 for b0, b1, _, _ in ob.ready)
    ws_bytes = _lib.lib().khmm_viterbi_workspace_bytes(bmax, ob.T, N)
    ws, wsp = workspace(ws_bytes, 256)      # cached: the backpointer workspace is 17 GB at config 3
    if p.kind == "gauss":
        # EXTENSION (include/khmm.h): log-density constants on the host, exactly as oracle.log_gaussian_pdf_1d forms them
        mean, sd = np.asarray(p.mean, dtype=np.float64), np.asarray(p.sd, dtype=np.float64)
        with np.errstate(divide="ignore"):
            logc = -np.log(sd) - np.log(np.sqrt(2.0 * np.pi))
        d_tab = _dev(np.stack([mean, sd, logc]), t.float64)
    else:
        with np.errstate(divide="ignore"):
            logBsm = np.log(p.B.T)
        d_logBsm = _dev(logBsm, t.float64)
    host_path = back = None
    if ob.ready is not None:
        try:
            host_path = t.empty((ob.B, ob.T), dtype=path_dtype, pin_memory=True)
            back = _side_streams.get(("d2h", dev.index))
            if back is None:
                back = _side_streams[("d2h", dev.index)] = t.cuda.Stream(device=dev)
        except RuntimeError:                # page-locked memory exhausted: results come back after the last block
            host_path = back = None
    main = t.cuda.current_stream()
    for k, blk in enumerate(blocks):
        b0, b1 = ob.wait_block(k) if blk is None else blk
        len_c = None if ob.lengths is None else ob.lengths[b0:b1]
        if p.kind == "gauss":
            _lib.call("khmm_viterbi_gauss_f64", b1 - b0, ob.T, N, ptr(d_logpi), ptr(d_logA), ptr(d_tab),
                      ptr(ob.tensor[b0:b1]), ob.code, ptr(len_c), wsp, ws_bytes, ptr(path[b0:b1]), _PATH_CODES[pname],
                      ptr(logp[b0:b1]), stream_ptr())
        else:
            _lib.call("khmm_viterbi_discrete_f64", b1 - b0, ob.T, N, p.M, ptr(d_logpi), ptr(d_logA), ptr(d_logBsm),
                      ptr(ob.tensor[b0:b1]), ob.code, ptr(len_c), wsp, ws_bytes, ptr(path[b0:b1]), _PATH_CODES[pname],
                      ptr(logp[b0:b1]), stream_ptr())
        if host_path is not None:
            back.wait_stream(main)
            with t.cuda.stream(back):
                host_path[b0:b1].copy_(path[b0:b1], non_blocking=True)
    if ob.ready is not None:
        ob.ready = None
        if host_path is not None:
            path.record_stream(back)
            host_logp = logp.cpu().numpy()          # synchronises the main stream
            back.synchronize()
            return HostResult(host_path.numpy()), HostResult(host_logp)
    return path, logp


class HostResult:
    """A result that is already on the host (block-wise paths of an overlapped Viterbi call); `finish` passes it on."""

    def __init__(self, array):
        self.array = array


def gamma(alpha, beta, scale, kind=_lib.GAMMA_REFERENCE):
    t = require_cuda()
    out = t.empty_like(alpha)
    suffix = "f32" if alpha.dtype == t.float32 else "f64"
    N = alpha.shape[-1]
    _lib.call("khmm_gamma_" + suffix, alpha.numel() // N, N, ptr(alpha), ptr(beta), ptr(scale), ptr(out), kind,
              stream_ptr())
    return out


def zi_single(A, E, alpha, beta):
    """Materialised zi for one sequence (float64 device tensors) -> (T-1, N, N)."""
    t = require_cuda()
    T, N = alpha.shape
    out = t.empty((max(T - 1, 0), N, N), dtype=t.float64, device=device())
    _lib.call("khmm_zi_f64", T, N, ptr(A), ptr(E), ptr(alpha), ptr(beta), ptr(out), stream_ptr())
    return out


# --------------------------------------------------------------------------- Baum-Welch
class Counts:
    """View helpers over the packed float64 count buffer (layout in include/khmm.h)."""

    def __init__(self, N, M, buf=None):
        t = require_cuda()
        self.N, self.M = N, M
        n = _lib.lib().khmm_counts_len(N, M)
        self.buf = buf if buf is not None else t.zeros((n,), dtype=t.float64, device=device())
        o = 0
        self.sl = {}
        for name, size in (("pi", N), ("xi", N * N), ("gam", N), ("emis_num", M * N), ("emis_den", N), ("tail", 3)):
            self.sl[name] = slice(o, o + size)
            o += size

    def host(self):
        h = self.buf.cpu().numpy()
        N, M = self.N, self.M
        return dict(pi=h[self.sl["pi"]], xi_raw=h[self.sl["xi"]].reshape(N, N), gam=h[self.sl["gam"]],
                    emis_num=h[self.sl["emis_num"]].reshape(M, N).T, emis_den=h[self.sl["emis_den"]],
                    loglik=h[self.sl["tail"]][0], nseq=h[self.sl["tail"]][1], flags=h[self.sl["tail"]][2])


def estep_chunk_size(B, T, N, esize, reserve=0.75):
    """How many sequences fit at once: alpha + V stashes [B][T][N] and scale [B][T]."""
    t = torch()
    free, _ = t.cuda.mem_get_info()
    per_seq = T * N * esize * 2 + T * esize + 64
    return int(max(1, min(B, (free * reserve) // per_seq)))


# rows (sequences x steps) from which `tensor_cores="auto"` picks the tf32 E-step: its per-term
# rounding error (~5e-4, unbiased) averages out as 1/sqrt(rows per count), see DESIGN.md 3.5
TC_AUTO_MIN_ROWS = 1 << 22


def estep_tc_supported(p: ModelParams, ob: Obs, real) -> bool:
    """The fused tcgen05 E-step covers 8 <= N <= 128 (N < 128 padded with unreachable states), float32, equal-length
    or ragged batches.  Measured at B = 131072, T = 512: 24 ms whatever N, against 132 / 157 / 153 / 222 / 467 ms on the
    float32 SIMT kernels at N = 8 / 16 / 32 / 64 / 100 (tools/dbg_estep_n.py).  Below 8 states the tf32 rounding of the
    transition matrix stops averaging out over a row (N = 3: log P off by 1.8e-5 relative, measured): those stay on the
    float32 kernels."""
    return (p.kind == "discrete" and 8 <= p.N <= 128 and real == torch().float32 and ob.T < 2 ** 31)


_ws_cache = {}


def _ws_key():
    """Scratch buffers are per (device, stream): kernels queued on different streams must not share scratch, and a
    block allocated under one stream is only ever reused by work on that stream (the caching allocator's contract)."""
    t = torch()
    return (t.cuda.current_device(), t.cuda.current_stream().cuda_stream)


def workspace(nbytes: int, align: int = 1024):
    """A cached, aligned device scratch buffer of at least `nbytes` (one per device and stream, grown on demand).
    The tensor-core E-step stash is tens of GB: re-allocating it every iteration costs a cudaMalloc of
    that size whenever the caching allocator has let the block go."""
    t = torch()
    dev = device()
    key = _ws_key()
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes + align:
        _ws_cache.pop(key, None)
        buf = None
        buf = t.empty((nbytes + align,), dtype=t.uint8, device=dev)
        _ws_cache[key] = buf
    return buf, buf.data_ptr() + (-buf.data_ptr()) % align


def release_workspace():
    """Drop the cached scratch buffers (they are otherwise kept for the life of the process)."""
    _ws_cache.clear()


def _pick_counts(big: "Counts", counts: "Counts", n: int):
    """Add the real states' statistics of a 128-state (padded) count buffer into an n-state one (device ops)."""
    b, c, M = big.buf, counts.buf, counts.M
    c[counts.sl["pi"]] += b[big.sl["pi"]][:n]
    c[counts.sl["xi"]] += b[big.sl["xi"]].view(128, 128)[:n, :n].reshape(-1)
    c[counts.sl["gam"]] += b[big.sl["gam"]][:n]
    c[counts.sl["emis_num"]] += b[big.sl["emis_num"]].view(M, 128)[:, :n].reshape(-1)
    c[counts.sl["emis_den"]] += b[big.sl["emis_den"]][:n]
    c[counts.sl["tail"]] += b[big.sl["tail"]]
    return counts


def estep_discrete_tc(p: ModelParams, ob: Obs, counts: "Counts" = None, chunk=None, stash_out=False, operands=None,
                      precision=None):
    """Fused tensor-core E-step (khmm_estep_tc_discrete_f32).  With `stash_out` also returns the
    forward stash W[b][t][:] = c_t alpha_t and the scale factors of the LAST chunk (tests).
    `operands`: device-resident float32 operand set at 128 states (pi, A, AT, Bsm attributes -- the training loop's,
    `DiscreteTrainer`); `p` is then only consulted for its shape (N may be < 128: the operands are already padded)."""
    t = require_cuda()
    if precision is None:
        precision = ESTEP_TC_DEFAULT_PRECISION
    if p.N < 128:
        # pad to 128 states with unreachable ones (pad_states): their statistics are exactly zero; the real states'
        # counts are then picked out of the 128-state buffer.  (The per-sequence smoothing of gamma_0 floors the padded
        # entries at 1e-10 like any other zero: a 1e-8 relative effect on pi, far below the tf32 bound.)
        if stash_out:
            raise ValueError("stash_out needs N = 128")
        n = p.N
        if operands is not None:
            big = estep_discrete_tc(ModelParams(kind="discrete", pi=np.empty(128), A=np.empty((128, 0)),
                                                B=np.empty((128, p.M))), ob, chunk=chunk, operands=operands,
                                    precision=precision)
        else:
            big = estep_discrete_tc(pad_states(p), ob, chunk=chunk, precision=precision)
        return _pick_counts(big, counts or Counts(n, p.M), n)
    dm = operands if operands is not None else DeviceModel(p, t.float32)
    N, M, dev = p.N, p.M, device()
    counts = counts or Counts(N, M)
    L = _lib.lib()
    if chunk is None:
        cached = _ws_cache.get(_ws_key())
        if cached is not None and cached.numel() >= L.khmm_estep_tc_workspace_bytes(ob.B, ob.T, N, M) + 1024:
            chunk = ob.B                      # the whole batch fits the scratch buffer we already hold
        else:
            free, _ = t.cuda.mem_get_info()
            if cached is not None:
                free += cached.numel()
            per_seq = ob.T * N * 4 + ob.T * 4 + 16
            chunk = int(max(128, min(ob.B, (free * 0.8) // per_seq - 128)))
            if chunk < ob.B:
                chunk -= chunk % 128
    nb0 = min(chunk, ob.B)
    nbytes = L.khmm_estep_tc_workspace_bytes(nb0, ob.T, N, M)
    ws, wsp = workspace(nbytes)
    W = sc = None
    if stash_out:
        W = t.zeros(((nb0 + 127) // 128, ob.T, 128, N), dtype=t.float32, device=dev)   # tile-major stash
        sc = t.zeros(((nb0 + 127) // 128, ob.T, 128), dtype=t.float32, device=dev)           # tile-major c
    # an overlapped upload is consumed block by block (each block waits for its own copy only); afterwards -- and for
    # every other batch -- the blocks are the chunks the scratch buffer allows
    blocks = [(0, ob.B)]
    if ob.ready is not None:
        if stash_out:
            ob.wait_all()
        else:
            blocks = [None] * len(ob.ready)
    for k, blk in enumerate(blocks):
        lo, hi = ob.wait_block(k) if blk is None else blk
        for b0 in range(lo, hi, chunk):
            nb = min(chunk, hi - b0)
            obs_c = ob.tensor[b0:b0 + nb]
            len_c = ob.lengths[b0:b0 + nb] if ob.lengths is not None else None
            _lib.call("khmm_estep_tc_discrete_ex_f32", nb, ob.T, N, M, ptr(dm.pi), ptr(dm.A), ptr(dm.AT), ptr(dm.Bsm),
                      ptr(obs_c), ob.code, ptr(len_c), wsp, nbytes, ptr(counts.buf), ptr(W), ptr(sc), int(precision),
                      stream_ptr())
    ob.ready = None
    if stash_out:
        W = W.permute(0, 2, 1, 3).reshape(-1, ob.T, N)[:nb0]     # -> [b][t][:]
        sc = sc.permute(0, 2, 1).reshape(-1, ob.T)[:nb0]
        return counts, W, sc
    return counts


def estep_uses_tc(p: ModelParams, ob: Obs, real, tensor_cores="auto") -> bool:
    use, _ = estep_precision(tensor_cores)
    if use == "auto":
        use = ob.B * ob.T >= TC_AUTO_MIN_ROWS
    return bool(use) and estep_tc_supported(p, ob, real)


def estep_discrete(p: ModelParams, ob: Obs, real, counts: Counts = None, chunk=None, tensor_cores="auto", operands=None):
    """Accumulate the reference-weighted Baum-Welch statistics of a batch into `counts`.
    tensor_cores: True / False / "auto" (tf32 tcgen05 path for N = 128 once the batch is large).
    `operands`: device-resident operand set in `real` (see `DiscreteTrainer`) used instead of uploading `p`."""
    t = require_cuda()
    if p.kind != "discrete":
        raise NotImplementedError("Baum-Welch is defined for discrete emissions "
                                  "(GaussianHMM.do_pass is broken in the reference, SURVEY.md Q9)")
    if estep_uses_tc(p, ob, real, tensor_cores):
        return estep_discrete_tc(p, ob, counts=counts, chunk=chunk, operands=operands,
                                 precision=estep_precision(tensor_cores)[1])
    ob.wait_all()
    dm = operands if operands is not None else DeviceModel(p, real)
    N, M, dev = p.N, p.M, device()
    counts = counts or Counts(N, M)
    esize = 4 if real == t.float32 else 8
    chunk = chunk or estep_chunk_size(ob.B, ob.T, N, esize)
    xi_ptr = counts.buf[counts.sl["xi"]].data_ptr()
    sfx = dm.suffix
    alpha = t.empty((min(chunk, ob.B), ob.T, N), dtype=real, device=dev)
    V = t.empty_like(alpha)
    scale = t.empty((min(chunk, ob.B), ob.T), dtype=real, device=dev)
    for b0 in range(0, ob.B, chunk):
        nb = min(chunk, ob.B - b0)
        obs_c = ob.tensor[b0:b0 + nb]
        len_c = None if ob.lengths is None else ob.lengths[b0:b0 + nb]
        _lib.call("khmm_forward_discrete_" + sfx, nb, ob.T, N, M, ptr(dm.pi), ptr(dm.A), ptr(dm.Bsm), ptr(obs_c),
                  ob.code, ptr(len_c), ptr(alpha), ptr(scale), None, stream_ptr())
        _lib.call("khmm_estep_discrete_" + sfx, nb, ob.T, N, M, ptr(dm.AT), ptr(dm.Bsm), ptr(obs_c), ob.code,
                  ptr(len_c), ptr(alpha), ptr(scale), ptr(V), ptr(counts.buf), stream_ptr())
        _lib.call("khmm_xi_accumulate_" + sfx, nb, ob.T, N, ptr(len_c), ptr(alpha), ptr(V), xi_ptr, stream_ptr())
    return counts


def mstep_discrete(p: ModelParams, counts: Counts):
    """Device M-step -> (pi, A, B) float64 NumPy; raises what the reference would raise."""
    t = require_cuda()
    N, M, dev = p.N, p.M, device()
    A_old = _dev(p.A, t.float64)
    pi_new = t.empty((N,), dtype=t.float64, device=dev)
    A_new = t.empty((N, N), dtype=t.float64, device=dev)
    B_new = t.empty((N, M), dtype=t.float64, device=dev)
    status = t.zeros((1,), dtype=t.int32, device=dev)
    _lib.call("khmm_mstep_discrete_f64", N, M, ptr(counts.buf), ptr(A_old), ptr(pi_new), ptr(A_new), ptr(B_new),
              ptr(status), stream_ptr())
    _raise_mstep_status(int(status.item()))
    return pi_new.cpu().numpy(), A_new.cpu().numpy(), B_new.cpu().numpy()


def _raise_mstep_status(st: int):
    if st & 1:
        raise ZeroDivisionError("float division by zero")          # util.py:57-58 with k == n
    if st & 2:
        raise ValueError("operands could not be broadcast together "
                         "(smooth_probabilities 2-D path with 2 <= k < n small entries, util.py:51-53)")
    if st & 4:
        raise AssertionError("sanity_check failed: a re-estimated distribution does not sum to 1")


class _Operands:
    """Device operand set of a discrete model in one real type: pi[Np], A[Np][Np], AT, symbol-major Bsm[M][Np]."""

    def __init__(self, Np, M, real):
        t, dev = torch(), device()
        self.real = real
        self.suffix = "f32" if real == t.float32 else "f64"
        self.pi = t.empty((Np,), dtype=real, device=dev)
        self.A = t.empty((Np, Np), dtype=real, device=dev)
        self.AT = t.empty((Np, Np), dtype=real, device=dev)
        self.Bsm = t.empty((M, Np), dtype=real, device=dev)


class DiscreteTrainer:
    """Device-resident Baum-Welch loop for a DiscreteHMM (AbstractHMM.train + DiscreteHMM.do_pass,
    abstract_hmm.py:400-441 / discrete_hmm.py:95-191, batched): the float64 master parameters, the operand copies the
    kernels read and the count buffer all stay on the GPU between iterations -- E-step, count all-reduce, M-step
    (`khmm_mstep_discrete_f64`) and the hand-over (`khmm_train_commit_discrete_*`) are queued on the stream without a
    host synchronisation; the host reads parameters, log-likelihood history and the reference's exceptional-case flags
    once, in `finish()`."""

    def __init__(self, p: ModelParams, real, max_iterations: int = 1):
        t = require_cuda()
        if p.kind != "discrete":
            raise NotImplementedError("Baum-Welch is defined for discrete emissions "
                                      "(GaussianHMM.do_pass is broken in the reference, SURVEY.md Q9)")
        dev = device()
        self.shape = ModelParams(kind="discrete", pi=np.empty(p.N), A=np.empty((p.N, 0)), B=np.empty((p.N, p.M)))
        self.N, self.M, self.real = p.N, p.M, real
        self.pi64, self.A64, self.B64 = _dev(p.pi, t.float64), _dev(p.A, t.float64), _dev(p.B, t.float64)
        self.pi_n, self.A_n, self.B_n = t.empty_like(self.pi64), t.empty_like(self.A64), t.empty_like(self.B64)
        self.status_it = t.zeros((1,), dtype=t.int32, device=dev)
        self.sticky = t.zeros((2,), dtype=t.int32, device=dev)
        self.history = t.zeros((max(1, max_iterations),), dtype=t.float64, device=dev)
        self.counts = Counts(self.N, self.M)
        self.ops = {}           # Np -> _Operands (N itself, or 128 for the padded tensor-core E-step)
        self.iteration = 0

    def operands(self, Np):
        ops = self.ops.get(Np)
        if ops is None:
            ops = self.ops[Np] = _Operands(Np, self.M, self.real)
            self._commit(ops, Np, None, initial=True)
        return ops

    def _commit(self, ops, Np, status, initial=False):
        src = (self.pi64, self.A64, self.B64) if initial else (self.pi_n, self.A_n, self.B_n)
        _lib.call("khmm_train_commit_discrete_" + ops.suffix, self.N, self.M, Np, self.iteration, ptr(status),
                  None if initial else ptr(self.sticky), ptr(src[0]), ptr(src[1]), ptr(src[2]), ptr(self.pi64),
                  ptr(self.A64), ptr(self.B64), ptr(ops.pi), ptr(ops.A), ptr(ops.AT), ptr(ops.Bsm), stream_ptr())

    def step(self, ob: Obs, chunk=None, tensor_cores="auto", group=None):
        """One Baum-Welch iteration over this rank's batch, asynchronous on the current stream."""
        from . import dist as kdist
        use_tc = estep_uses_tc(self.shape, ob, self.real, tensor_cores)
        Np = 128 if use_tc else self.N
        ops = self.operands(Np)
        self.counts.buf.zero_()
        estep_discrete(self.shape, ob, self.real, counts=self.counts, chunk=chunk,
                       tensor_cores=(tensor_cores if (use_tc and isinstance(tensor_cores, str) and tensor_cores != "auto")
                                     else use_tc), operands=ops)
        kdist.all_reduce_counts(self.counts.buf, group=group)
        if self.iteration < self.history.numel():
            o = self.counts.sl["tail"].start
            self.history[self.iteration:self.iteration + 1].copy_(self.counts.buf[o:o + 1])
        _lib.call("khmm_mstep_discrete_f64", self.N, self.M, ptr(self.counts.buf), ptr(self.A64), ptr(self.pi_n),
                  ptr(self.A_n), ptr(self.B_n), ptr(self.status_it), stream_ptr())
        # every operand set in use follows the masters (normally one)
        first = True
        for np_, o_ in self.ops.items():
            if first:
                self._commit(o_, np_, self.status_it)
                first = False
            else:
                self._commit(o_, np_, None, initial=True)
        self.iteration += 1

    def last_loglik(self) -> float:
        """Summed log-likelihood (under the parameters the iteration started from) of the last iteration; synchronises."""
        return float(self.history[self.iteration - 1].item())

    def finish(self):
        """-> (pi, A, B (N, M), loglik history) on the host; raises what the reference's do_pass would have raised in
        the first iteration that hit an exceptional case (the parameters are those of the last good iteration)."""
        t = torch()
        sticky = self.sticky.cpu().numpy()
        hist = self.history[:self.iteration].cpu().numpy().copy()
        pi, A, B = self.pi64.cpu().numpy(), self.A64.cpu().numpy(), self.B64.cpu().numpy()
        self.failed_iteration = int(sticky[1]) - 1 if sticky[0] else None
        self.status = int(sticky[0])
        return pi, A, B, hist


# ---------------------------------------------------------------------------------- sampling
def _cdf(p):
    """The CDF numpy.random.choice(p=...) searches: cumsum(p) / cumsum(p)[-1] (last axis)."""
    c = np.cumsum(np.asarray(p, dtype=np.float64), axis=-1)
    return c / c[..., -1:]


def simulate(p: ModelParams, B: int, T: int, seed: int = 0, uniforms=None):
    """Batched AbstractHMM.simulate on the device -> (states[B][T] int32, emissions[B][T]) device tensors
    (int32 symbols for discrete models, float64 values for 1-D Gaussians)."""
    t = require_cuda()
    dev = device()
    if B < 0 or T < 1:
        raise ValueError("simulate needs B >= 0 and T >= 1")
    cum_pi, cumA = _dev(_cdf(p.pi), t.float64), _dev(_cdf(p.A), t.float64)
    U = None
    if uniforms is not None:
        U = t.as_tensor(np.ascontiguousarray(uniforms, dtype=np.float64)).to(dev)
        if tuple(U.shape) != (B, T, 3):
            raise ValueError("uniforms must have shape (B, T, 3)")
    states = t.empty((B, T), dtype=t.int32, device=dev)
    if p.kind == "discrete":
        cumB = _dev(_cdf(p.B), t.float64)
        out = t.empty((B, T), dtype=t.int32, device=dev)
        _lib.call("khmm_simulate_discrete", B, T, p.N, p.M, ptr(cum_pi), ptr(cumA), ptr(cumB), int(seed) & (2 ** 64 - 1),
                  ptr(U), ptr(states), ptr(out), stream_ptr())
    elif p.kind == "gauss":
        mean, sd = _dev(p.mean, t.float64), _dev(p.sd, t.float64)
        out = t.empty((B, T), dtype=t.float64, device=dev)
        _lib.call("khmm_simulate_gauss", B, T, p.N, ptr(cum_pi), ptr(cumA), ptr(mean), ptr(sd), int(seed) & (2 ** 64 - 1),
                  ptr(U), ptr(states), ptr(out), stream_ptr())
    else:
        raise NotImplementedError("batched sampling covers discrete and 1-D Gaussian emissions")
    return states, out


# ------------------------------------------------------------- Gaussian Baum-Welch (extension)
class GaussCounts:
    """View helpers over the packed float64 count buffer of the Gaussian E-step (include/khmm.h)."""

    def __init__(self, N, buf=None):
        t = require_cuda()
        self.N = N
        n = _lib.lib().khmm_counts_len_gauss(N)
        self.buf = buf if buf is not None else t.zeros((n,), dtype=t.float64, device=device())
        o = 0
        self.sl = {}
        for name, size in (("pi", N), ("xi", N * N), ("gam", N), ("s0", N), ("s1", N), ("s2", N), ("tail", 3)):
            self.sl[name] = slice(o, o + size)
            o += size

    def host(self):
        h = self.buf.cpu().numpy()
        N = self.N
        return dict(pi=h[self.sl["pi"]], xi_raw=h[self.sl["xi"]].reshape(N, N), gam=h[self.sl["gam"]],
                    s0=h[self.sl["s0"]], s1=h[self.sl["s1"]], s2=h[self.sl["s2"]],
                    loglik=h[self.sl["tail"]][0], nseq=h[self.sl["tail"]][1], flags=h[self.sl["tail"]][2])


def estep_gauss(p: ModelParams, ob: Obs, real, counts: GaussCounts = None, chunk=None):
    """Textbook-EM statistics of a batch for 1-D Gaussian emissions (extension, include/khmm.h)."""
    t = require_cuda()
    if p.kind != "gauss":
        raise NotImplementedError("Gaussian Baum-Welch covers 1-D Gaussian emissions")
    dm = DeviceModel(p, real)
    N, dev = p.N, device()
    counts = counts or GaussCounts(N)
    esize = 4 if real == t.float32 else 8
    if chunk is None:
        free, _ = t.cuda.mem_get_info()
        chunk = int(max(1, min(ob.B, (free * 0.7) // (ob.T * N * esize * 3 + ob.T * esize + 64))))
    sfx = dm.suffix
    nb0 = min(chunk, ob.B)
    alpha = t.empty((nb0, ob.T, N), dtype=real, device=dev)
    beta = t.empty_like(alpha)
    V = t.empty_like(alpha)
    scale = t.empty((nb0, ob.T), dtype=real, device=dev)
    xi_ptr = counts.buf[counts.sl["xi"]].data_ptr()
    for b0 in range(0, ob.B, chunk):
        nb = min(chunk, ob.B - b0)
        obs_c = ob.tensor[b0:b0 + nb]
        len_c = None if ob.lengths is None else ob.lengths[b0:b0 + nb]
        _lib.call("khmm_forward_gauss_" + sfx, nb, ob.T, N, ptr(dm.pi), ptr(dm.A), ptr(dm.mean), ptr(dm.sd), ptr(obs_c),
                  ob.code, ptr(len_c), ptr(alpha), ptr(scale), None, stream_ptr())
        _lib.call("khmm_backward_gauss_" + sfx, nb, ob.T, N, ptr(dm.AT), ptr(dm.mean), ptr(dm.sd), ptr(obs_c), ob.code,
                  ptr(len_c), ptr(scale), None, ptr(beta), None, _lib.GAMMA_POSTERIOR, stream_ptr())
        _lib.call("khmm_estep_gauss_" + sfx, nb, ob.T, N, ptr(dm.mean), ptr(dm.sd), ptr(obs_c), ob.code, ptr(len_c),
                  ptr(alpha), ptr(beta), ptr(scale), ptr(V), ptr(counts.buf), stream_ptr())
        _lib.call("khmm_xi_accumulate_" + sfx, nb, ob.T, N, ptr(len_c), ptr(alpha), ptr(V), xi_ptr, stream_ptr())
    return counts


def mstep_gauss(p: ModelParams, counts: GaussCounts, sd_min: float = 1e-6):
    """Device M-step -> (pi, A, mean, sd) float64 NumPy."""
    t = require_cuda()
    N, dev = p.N, device()
    A_old = _dev(p.A, t.float64)
    pi_new = t.empty((N,), dtype=t.float64, device=dev)
    A_new = t.empty((N, N), dtype=t.float64, device=dev)
    mean_new = t.empty((N,), dtype=t.float64, device=dev)
    sd_new = t.empty((N,), dtype=t.float64, device=dev)
    status = t.zeros((1,), dtype=t.int32, device=dev)
    _lib.call("khmm_mstep_gauss_f64", N, ptr(counts.buf), ptr(A_old), float(sd_min), ptr(pi_new), ptr(A_new),
              ptr(mean_new), ptr(sd_new), ptr(status), stream_ptr())
    st = int(status.item())
    if st & 1:
        raise ZeroDivisionError("float division by zero")
    if st & 2:
        raise ValueError("operands could not be broadcast together "
                         "(smooth_probabilities 2-D path with 2 <= k < n small entries, util.py:51-53)")
    if st & 8:
        raise ValueError("a state received no posterior mass: its mean and variance are undefined")
    if st & 4:
        raise AssertionError("sanity_check failed: a re-estimated distribution does not sum to 1")
    return pi_new.cpu().numpy(), A_new.cpu().numpy(), mean_new.cpu().numpy(), sd_new.cpu().numpy()
