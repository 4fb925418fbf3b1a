#!/usr/bin/env python
"""Benchmark of the batched-HMM hot path (BASELINE.json metric: state-transitions/s = B*T*N^2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

Default workload = BASELINE.json configs[1] ("c2"): GaussianHMM, 8 states, forward-backward
log-likelihood of 16,384 synthetic sequences of T=2,048 per GPU (weak scaling: every rank owns its
own 16,384 sequences; the path shards by sequence with no data-path collective).  One "step" =
one pass of the hot path over the batch.  Prints ONE JSON line (rank 0).

  value     device-resident throughput: inputs already in HBM, CUDA events around K steps
  e2e       the same metric through the public API (GaussianHMM.log_likelihood_batch) with pinned
            HOST buffers: H2D of the observations + kernels + D2H of the log-likelihoods per step
  roofline  dominant kernel against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle (NumPy port of the reference algorithm) timed on the host cores on a
            bounded sample -- a reported baseline only
`--impl reference` times that CPU implementation alone with all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, N, M, B per GPU, T)
    "c1": ("DiscreteHMM N=3 M=4, forward + backward + Viterbi of ONE sequence, T=1000 (the reference's CPU-runnable case; "
           "latency, not throughput)", 3, 4, 1, 1000),
    "c2": ("GaussianHMM N=8 D=1 forward-backward log-likelihood, B=16384 T=2048 per GPU", 8, 0, 16384, 2048),
    "c3": ("DiscreteHMM N=64 M=256 Viterbi, B=65536 T=4096", 64, 256, 65536, 4096),
    "c4": ("DiscreteHMM N=128 M=1024 Baum-Welch iteration, B=131072 T=512 per GPU", 128, 1024, 131072, 512),
    "c5": ("GaussianHMM N=1024 D=1 forward-backward, B=4096 T=1024", 1024, 0, 4096, 1024),
}
CFG_INDEX = {"c1": 1, "c2": 2, "c3": 3, "c4": 4, "c5": 5}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------- synthetic data
def make_gaussian(N, B, T, seed):
    import kerehmm_b200 as K
    np.random.seed(seed)
    model = K.GaussianHMM(N, 1, random_transitions=True, random_emissions=True)
    mean = np.array([d.mean for d in model.emissionDistributions])
    rng = np.random.default_rng(1000 + seed)
    obs = np.empty((B, T), dtype=np.float32)
    step = max(1, (1 << 24) // T)
    for b0 in range(0, B, step):
        nb = min(step, B - b0)
        obs[b0:b0 + nb] = mean[rng.integers(0, N, size=(nb, T))] + rng.standard_normal((nb, T), dtype=np.float32)
    return model, obs


def make_discrete(N, M, B, T, seed):
    import kerehmm_b200 as K
    np.random.seed(seed)
    model = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
    rng = np.random.default_rng(1000 + seed)
    dt = np.uint8 if M <= 256 else np.uint16
    obs = rng.integers(0, M, size=(B, T), dtype=dt)
    return model, obs


# ------------------------------------------------------------------------------ CPU legs
def _cpu_chunk(args):
    kind, params, obs = args
    from oracle import hmm_oracle as O
    t0 = time.perf_counter()
    if kind == "c2" or kind == "c5":
        pi, A, mean, sd = params
        E = O.emissions_gaussian_1d(mean, sd, obs.astype(np.float64))
        O.fwdbwd_loglik_batch(pi, A, E)
    elif kind == "c1":
        pi, A, Bm = params
        o = obs.astype(np.int64)
        E = O.emissions_discrete(Bm, o)
        for b in range(o.shape[0]):
            alpha, scale = O.forward(pi, A, E[b])
            O.backward(A, E[b], scale)
        with np.errstate(divide="ignore"):
            O.viterbi(pi, A, Bm, o)
    elif kind == "c3":
        pi, A, Bm = params
        with np.errstate(divide="ignore"):
            O.viterbi(pi, A, Bm, obs.astype(np.int64))
    elif kind == "c4":
        pi, A, Bm = params
        O.mstep_discrete(O.batch_stats_discrete(pi, A, Bm, obs.astype(np.int64)))
    return time.perf_counter() - t0


# bounded samples: a few seconds of wall time of the NumPy port on the GPU box's 16+ host cores each
CPU_SAMPLE = {"c1": (1, 1000), "c2": (16384, 2048), "c3": (8192, 1024), "c4": (8192, 512), "c5": (32, 32)}


def cpu_params(kind, model):
    pi, A = model.initialProbabilities.copy(), model.transitionMatrix.copy()
    if isinstance(model, _HostModel):
        return (pi, A, model._mean, model._sd) if kind in ("c2", "c5") else (pi, A, model._B)
    if kind in ("c2", "c5"):
        return (pi, A, np.array([d.mean for d in model.emissionDistributions], dtype=np.float64),
                np.array([d.variance for d in model.emissionDistributions], dtype=np.float64))
    return (pi, A, model.emission_matrix())


def cpu_leg(kind, model, obs, cores):
    """Oracle port on a bounded sample split over `cores` worker processes -> (transitions/s, sample text)."""
    import multiprocessing as mp
    N = model.nStates
    sb, st = CPU_SAMPLE[kind]
    sb, st = min(sb, obs.shape[0]), min(st, obs.shape[1])
    sample = np.ascontiguousarray(obs[:sb, :st])
    params = cpu_params(kind, model)
    cores = max(1, min(cores, sb))
    chunks = [(kind, params, c) for c in np.array_split(sample, cores) if len(c)]
    t0 = time.perf_counter()
    if len(chunks) == 1:
        _cpu_chunk(chunks[0])
    else:
        with mp.get_context("fork").Pool(len(chunks)) as pool:
            pool.map(_cpu_chunk, chunks)
    dt = time.perf_counter() - t0
    return sb * st * N * N / dt, "oracle port (NumPy fp64), B=%d T=%d at the workload's N, %d processes" % (
        sb, st, len(chunks)), len(chunks), dt


# ------------------------------------------------------------------------ clock sampling
class ClockSampler:
    """nvidia-smi polled every 100 ms.  Started BEFORE the warm-up (its start-up takes driver locks and a few
    hundred ms); only the samples whose timestamp falls inside the timed region are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self, t_begin=None, t_end=None):
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        rows = []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(c[1]), float(c[2]), float(c[3]), c[5:9]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        sel = rows
        if t_begin is not None:
            for slack in (0.0, 0.12, 0.3):      # short timed regions: take the polls nearest to them
                sel = [r for r in rows if t_begin - slack <= r[0] <= t_end + slack]
                if sel:
                    break
            else:
                sel = rows
        sm, mx, reasons, power = [], [], set(), []
        for _, a, b, c3, flags in sel:
            sm.append(a); mx.append(b); power.append(c3)
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), flags):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []          # the upper half = samples under load
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import kerehmm_b200 as K
    from kerehmm_b200 import _lib, engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    engine.require_cuda()
    kind = args.workload
    desc, N, M, B, T = WORKLOADS[kind]
    if args.batch:
        B = args.batch
    seed = CFG_INDEX[kind]
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    hbm_peak, peak_src = peaks()

    launches_per_step = 0
    ncu_traffic = None
    if kind in ("c2",):
        model, obs_h = make_gaussian(N, B, T, seed)
        if rank:      # different sequences per rank, same model
            obs_h = make_gaussian(N, B, T, seed + 100 * rank)[1]
        p = model._params()
        dm = engine.DeviceModel(p, torch.float32)
        host_sets = [torch.from_numpy(obs_h).pin_memory(), torch.from_numpy(obs_h[::-1].copy()).pin_memory()]
        dev_sets = [h.to(dev) for h in host_sets]
        ll = torch.empty((B,), dtype=torch.float64, device=dev)
        llb = torch.empty((B,), dtype=torch.float64, device=dev)

        def step(i):
            o = dev_sets[i & 1]
            _lib.call("khmm_fwdbwd_loglik_gauss_f32", B, T, N, engine.ptr(dm.pi), engine.ptr(dm.A),
                      engine.ptr(dm.mean), engine.ptr(dm.sd), engine.ptr(o), _lib.F32, None, engine.ptr(ll),
                      engine.ptr(llb), stream.cuda_stream)

        def e2e_step(i):
            out = model.log_likelihood_batch(host_sets[i & 1])
            return out[0]

        launches_per_step = 1
        algo_bytes = B * T * 4 + 2 * B * 8          # observations once + two float64 results per sequence
        h2d, d2h = B * T * 4, 2 * B * 8
        dominant = "fwdbwd_small_kernel<8, gauss>"
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full
        # (profiles/r1_ncu_c2_fwdbwd_small_summary.csv): the two directions share most lines in L2
        ncu_traffic = 180.484352e6 + 4.261632e6
        l2_note = "two alternating 134 MB observation sets (268 MB > 126 MB L2)"
        dtype = "f32"
    elif kind == "c1":
        model, obs_h = make_discrete(N, M, B, T, seed)
        p = model._params()
        ob = engine.prepare_obs(torch.from_numpy(obs_h.astype(np.int64)).to(dev), symbols=True)

        def step(i):          # the three single-sequence calls of the reference's tests, inputs resident on the device
            _, scale, _ = engine.forward(p, ob, torch.float64)
            engine.backward(p, ob, torch.float64, scale)
            engine.viterbi(p, ob)

        def e2e_step(i):      # ... and through the reference-named methods, NumPy in / NumPy out
            o = obs_h[0].astype(np.int64)
            _, scale = model.forward(o)
            model.backward(o, scale)
            return model.viterbi_path(o)[1]

        launches_per_step = 4      # forward, backward, Viterbi DP, backtrace
        algo_bytes = B * T * (8 + 3 * N * 8 + N + 8)     # symbols, alpha/scale/beta rows out, backpointers, path
        h2d, d2h = 3 * B * T * 8, B * T * (2 * N * 8 + 8 + 8)
        dominant = "forward/backward/viterbi generic kernels, one CTA: a 1000-step dependency chain (latency, no roofline claim)"
        l2_note = "9e3 transitions per step: everything is cache resident; the step is 4 launches of ~1000 dependent steps"
        dtype = "f64"
    elif kind == "c3":
        model, obs_h = make_discrete(N, M, B, T, seed)
        host_sets = [torch.from_numpy(obs_h).pin_memory()]
        ob = engine.prepare_obs(host_sets[0].to(dev), symbols=True)
        p = model._params()

        def step(i):
            engine.viterbi(p, ob, torch.uint8)

        def e2e_step(i):
            return model.viterbi_batch(host_sets[0], path_dtype="uint8")[1]

        launches_per_step = 2
        algo_bytes = B * T * (1 + N + 1 + 1)         # obs + N backpointer bytes written + 1 read + path byte
        h2d, d2h = B * T, B * T + B * 8
        dominant = "viterbi_screen_kernel<64, uint8> (float32 screen, float64 resolve) + viterbi_backtrace_kernel"
        # dram__bytes_read.sum + dram__bytes_write.sum of the DP kernel at full size, ncu --set full
        # (profiles/r1_ncu_c3_viterbi_screen_fullsize_summary.csv): 17.18 GB of backpointers written once, observations read once
        ncu_traffic = 0.309431808e9 + 17.133593e9
        l2_note = "backpointer workspace 17 GB per step (>> L2)"
        dtype = "f64"
    elif kind == "c4":
        model, obs_h = make_discrete(N, M, B, T, seed + 100 * rank)
        model = make_discrete(N, M, 1, 1, seed)[0]
        host_sets = [torch.from_numpy(obs_h.view(np.int16)).pin_memory()]
        ob = engine.prepare_obs(host_sets[0].to(dev), symbols=True, u16_bits=True)

        import ctypes
        kernel_ms = []          # (forward, backward) device durations of every step, CUDA events inside libkhmm

        def step(i):
            from kerehmm_b200 import dist as kdist
            p_ = model._params()
            counts = engine.estep_discrete(p_, ob, torch.float32)
            kdist.all_reduce_counts(counts.buf)
            model._assign(*engine.mstep_discrete(p_, counts))
            f, b = ctypes.c_float(), ctypes.c_float()
            _lib.call("khmm_estep_tc_last_kernel_ms", ctypes.byref(f), ctypes.byref(b))
            kernel_ms.append((f.value, b.value))

        def e2e_step(i):
            o = engine.prepare_obs(host_sets[0], symbols=True, u16_bits=True)
            p_ = model._params()
            counts = engine.estep_discrete(p_, o, torch.float32)
            from kerehmm_b200 import dist as kdist
            kdist.all_reduce_counts(counts.buf)
            model._assign(*engine.mstep_discrete(p_, counts))
            return counts.buf[-3:].cpu()

        launches_per_step = 6      # 2 x bw_round_matrix, bw_forward_tc, bw_backward_tc, bw_fold, mstep_discrete
        # dominant kernel = bw_backward_tc_kernel: reads the forward stash once (512 B per (sequence, step)), two
        # scale factors and the 2-byte symbol (twice: epilogue + reduction warps)
        algo_bytes = B * T * (N * 4 + 8 + 4)
        h2d, d2h = B * T * 2, 24
        dominant = "bw_backward_tc_kernel (tcgen05 kind::tf32, xi in TMEM, L2 count reductions)"
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch at B=18944 (ncu --set full,
        # profiles/r1_ncu_c4_bw_backward_tc_summary.csv), scaled to this batch: traffic ~ algorithmic bytes
        ncu_traffic = (5.037033e9 + 0.004398e9) * (B * T) / (18944.0 * 512.0)
        l2_note = "forward stash of B*T*512 B per step (34 GB at B=131072, >> L2)"
        dtype = "tf32 multiply / f32 accumulate"
    elif kind == "c5":
        Bg = B // world if args.strong else B
        model, obs_h = make_gaussian(N, Bg, T, seed + 100 * rank)
        model = make_gaussian(N, 1, 1, seed)[0]
        host_sets = [torch.from_numpy(obs_h).pin_memory()]
        ob = engine.prepare_obs(host_sets[0].to(dev), symbols=False)
        p = model._params()
        B = Bg
        dm = engine.DeviceModel(p, torch.float32)
        ll = torch.empty((B,), dtype=torch.float64, device=dev)
        llb = torch.empty((B,), dtype=torch.float64, device=dev)
        c5_prec = engine.TC_BF16 if args.operands == "bf16" else engine.TC_TF32
        nbytes = _lib.lib().khmm_fwdbwd_loglik_tc_workspace_bytes(B, N)
        ws = torch.empty((nbytes + 256,), dtype=torch.uint8, device=dev)
        wsp = ws.data_ptr() + (-ws.data_ptr()) % 256

        def step(i):
            _lib.call("khmm_fwdbwd_loglik_tc_gauss_f32", B, T, N, engine.ptr(dm.pi), engine.ptr(dm.A),
                      engine.ptr(dm.AT), engine.ptr(dm.mean), engine.ptr(dm.sd), engine.ptr(ob.tensor), ob.code, wsp,
                      nbytes, engine.ptr(ll), engine.ptr(llb), c5_prec, stream.cuda_stream)

        def e2e_step(i):
            return model.log_likelihood_batch(host_sets[0], tensor_cores="bf16" if c5_prec else True)[0]

        launches_per_step = 5      # 2 x matrix conversion, tc_init, tc_recursion2 (all T-1 steps), tc_final
        algo_bytes = None
        h2d, d2h = B * T * 4, 2 * B * 8
        dominant = "tc_recursion2%s_kernel (tcgen05 kind::%s, one cooperative launch for all steps)" % (
            "" if c5_prec else "_pair", "f16 bf16 operands" if c5_prec else "tf32, cta_group::2")
        l2_note = "W ping-pong rows + A/AT stay L2 resident by design"
        dtype = ("bf16" if c5_prec else "tf32") + " multiply / f32 accumulate"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(args.warmup):
        step(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    ev0.record(stream)
    for i in range(args.steps):
        step(i)
    ev1.record(stream)
    barrier()
    wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(wall0, wall1) if sampler else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    ms_per_step = ms / args.steps
    transitions = float(B) * T * N * N * world
    value = transitions / (ms_per_step * 1e-3)

    # ---- end-to-end timing through the public API with host buffers
    e2e_steps = max(1, min(args.steps, 5))
    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        r = e2e_step(i)
        _ = r if isinstance(r, np.ndarray) else r
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = transitions / float(t_e.item())
    # the PCIe floor of the end-to-end number: the same input bytes, pinned host -> device, nothing else
    h2d_copy_ms = None
    if kind != "c1":
        src = host_sets[0]
        dst = torch.empty_like(src, device=dev)
        best = None
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        h2d_copy_ms = best * 1e3
        del dst

    out = None
    if rank == 0:
        if kind == "c5":
            tflops = 4.0 * transitions / world / (ms_per_step * 1e-3) / 1e12   # 2 flop fwd + 2 flop bwd per transition
            bf = args.operands == "bf16"
            peak = 1621.0 if bf else 1621.0 / 2
            roof = {"bound": "tensor", "achieved": tflops, "peak": peak, "unit": "TFLOP/s",
                    "frac": tflops / peak, "traffic": 43.3344e6 if bf else None, "kernel": dominant,
                    "traffic_note": "dram__bytes_read+write of one launch (profiles/r1_ncu_c5_tc_recursion2_bf16_summary.csv): "
                                    "W ping-pong rows and A stay in L2; L2->SM operand traffic is 201 GB per launch",
                    "peak_source": "measured bf16 burst peak (MEASURED_PEAKS.json)" if bf else
                    "half of the measured bf16 burst peak (MEASURED_PEAKS.json): tf32 runs at half the bf16 rate",
                    "accuracy": "log-likelihood vs float64 oracle at this T: max rel 6.8e-6 (bf16) / 9.2e-7 (tf32), measured"}
        elif kind == "c4":
            fwd_ms = float(np.mean([k[0] for k in kernel_ms[-args.steps:]]))
            bwd_ms = float(np.mean([k[1] for k in kernel_ms[-args.steps:]]))
            achieved = algo_bytes / (bwd_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": ncu_traffic, "kernel": dominant, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": bwd_ms,
                    "forward_kernel": {"name": "bw_forward_tc_kernel", "kernel_ms": fwd_ms,
                                       "achieved_gbs": B * T * (N * 4 + 4 + 2) / (fwd_ms * 1e-3) / 1e9},
                    "tensor": {"achieved_tflops": 4.0 * B * T * N * N / (bwd_ms * 1e-3) / 1e12, "peak_tflops": 1621.0 / 2,
                               "note": "MMA1 + MMA2 of the backward kernel; tf32 peak = half the measured bf16 burst peak"},
                    "note": "the kernel is bound by L2 reduction (red.global) throughput and LSU occupancy, "
                            "not by HBM or the tensor pipe (DESIGN.md 3.5)"}
        else:
            achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": ncu_traffic, "kernel": dominant, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": algo_bytes}
        out = {
            "metric": "state-transitions/sec (B*T*N^2)", "value": value, "unit": "transitions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if (kind == "c5" and args.strong) else "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic (seeded; reference-compatible random model init)",
            "config": {"workload": desc, "name": kind, "N": N, "M": M, "B_per_gpu": B, "T": T,
                       "sharding": "sequences sharded across ranks, no data-path collective" if kind != "c4"
                       else "sequences sharded across ranks, one count all-reduce per iteration",
                       "l2": l2_note},
            "e2e": {"value": e2e_value, "unit": "transitions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(t_e.item()) * 1e3, "api": "kerehmm_b200 public class API, pinned host buffers",
                    "h2d_copy_alone_ms": h2d_copy_ms},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof,
            "clocks": clocks,
        }
        if kind == "c3":
            # float32-screened kernel: 403 warp instructions per (sequence, step) = 3.15 per 32 transitions
            # (ncu smsp__inst_executed, profiles/r1_ncu_c3_viterbi_screen_summary.csv); a scheduler issues one warp
            # instruction per clock, 4 schedulers per SM.  The float64 register-tile kernel it replaces needs 2 FP64-pipe
            # instructions per transition (measured DADD rate 18.4e12/s -> 9.2e12 transitions/s, profiles/r1_microbench_pipes.json).
            ach = transitions / world / (ms_per_step * 1e-3)
            issue_bound = 148 * 4 * 1.965e9 * 32 / 3.15
            out["roofline"]["issue_slots"] = {"achieved_transitions_per_s": ach, "issue_bound_transitions_per_s": issue_bound,
                                              "frac": ach / issue_bound, "warp_instructions_per_32_transitions": 3.15,
                                              "fp64_pipe_bound_of_the_unscreened_kernel": 18.4e12 / 2,
                                              "note": "bit-exact float64 Viterbi with a float32 screen: instruction-issue "
                                                      "bound (2 warps per scheduler, 221 registers), not HBM bound; the "
                                                      "backpointer stream is 1.4 % of HBM (DESIGN.md 3.2)"}
        if kind in ("c2",):
            flops = transitions / world * 4.0          # 2 (forward) + 2 (backward) flop per transition
            ach_tf = flops / (ms_per_step * 1e-3) / 1e12
            out["roofline"]["fp32_simt"] = {"achieved_tflops": ach_tf, "nominal_peak_tflops": 74.4,
                                            "measured_packed_ffma2_peak_tflops": 67.7, "frac_of_measured_peak": ach_tf / 67.7,
                                            "fma_pipe_busy_ncu": 0.61,
                                            "note": "the binding roof of this kernel: FP32 issue, not HBM (DESIGN.md 3.1).  "
                                                    "achieved = the 4 N^2 flop per (sequence, step) of the two matrix-vector "
                                                    "products alone; with the emission and scaling arithmetic the FMA pipe "
                                                    "is 61 % busy (ncu, profiles/r1_ncu_c2_fwdbwd_small_summary.csv); the "
                                                    "packed-FFMA2 peak was measured on this pool "
                                                    "(profiles/r1_microbench_pipes.json)"}
        if world == 1 and not args.no_cpu:
            cv, sample, cores, dt = cpu_leg(kind, model, obs_h if kind != "c4" else obs_h, args.cpu_cores)
            out["cpu_baseline"] = {"value": cv, "unit": "transitions/s", "cores": cores, "kind": "port",
                                   "sample": sample, "seconds": dt}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return out


def run_reference(args):
    """The reference's algorithm on the host cores (oracle port; the reference itself is Python-2
    source that cannot be shipped or imported on the box -- DESIGN.md)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = args.workload
    desc, N, M, B, T = WORKLOADS[kind]
    seed = CFG_INDEX[kind]
    sb, st = CPU_SAMPLE[kind]
    if kind in ("c2", "c5"):
        model, obs = make_gaussian_host(N, sb, st, seed)
    else:
        model, obs = make_discrete_host(N, M, sb, st, seed)
    cores = args.cpu_cores
    for _ in range(args.warmup):
        cpu_leg(kind, model, obs, cores)
    vals, secs = [], []
    for _ in range(args.steps):
        v, sample, used, dt = cpu_leg(kind, model, obs, cores)
        vals.append(v); secs.append(dt)
    value = float(np.mean(vals))
    out = {"impl": "reference", "metric": "state-transitions/sec (B*T*N^2)", "value": value, "unit": "transitions/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic (seeded; same generator as the GPU arm)",
           "config": {"workload": desc, "name": kind, "N": N, "M": M, "B_per_gpu": B, "T": T},
           "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": used, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


class _HostModel:
    """Parameter holder for the CPU arm: built from the oracle's generators (same RNG draw order as
    the package constructors, tests/test_oracle_golden.py), so no product code is on this path."""

    def __init__(self, pi, A, mean=None, sd=None, Bm=None):
        self.initialProbabilities, self.transitionMatrix, self.nStates = pi, A, A.shape[0]
        self._mean, self._sd, self._B = mean, sd, Bm


def make_gaussian_host(N, B, T, seed):
    from oracle import hmm_oracle as O
    np.random.seed(seed)
    pi, A, mean, sd = O.random_gaussian_model(N)
    rng = np.random.default_rng(1000 + seed)
    obs = (mean[rng.integers(0, N, size=(B, T))] + rng.standard_normal((B, T), dtype=np.float32)).astype(np.float32)
    return _HostModel(pi, A, mean=mean, sd=sd), obs


def make_discrete_host(N, M, B, T, seed):
    from oracle import hmm_oracle as O
    np.random.seed(seed)
    pi, A, Bm = O.random_discrete_model(N, M)
    rng = np.random.default_rng(1000 + seed)
    obs = rng.integers(0, M, size=(B, T), dtype=np.uint8 if M <= 256 else np.uint16)
    return _HostModel(pi, A, Bm=Bm), obs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override sequences per GPU (debug)")
    ap.add_argument("--strong", action="store_true", help="c5: split the fixed batch across ranks")
    ap.add_argument("--operands", default="bf16", choices=["tf32", "bf16"],
                    help="c5: tensor-core operand precision (bf16: log-likelihoods within 7e-6 of float64 at c5's T=1024, "
                         "measured; tf32: 9e-7)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-cores", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
