#!/usr/bin/env python
"""Benchmark of the batched-HMM hot path (BASELINE.json metric: state-transitions/s = B*T*N^2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--also c2,c3,c5|none]
                    [--impl reference]

Default workload = BASELINE.json configs[3] ("c4"), the configuration north_star's two numeric targets are quoted
on: DiscreteHMM, 128 states x 1024 symbols, batched Baum-Welch on 131,072 sequences of T=512 PER GPU (1M sequences
on 8 GPUs).  One "step" = one Baum-Welch iteration: fused tensor-core E-step over this rank's shard, ONE NCCL sum
all-reduce of the packed float64 count buffer (libkhmm's own communicator, khmm_allreduce_counts), the device M-step
and the parameter hand-over to the next iteration -- all queued on one stream without a host synchronisation
(`engine.DiscreteTrainer`, what `DiscreteHMM.train_batch` runs).  Prints ONE JSON line (rank 0).

  value     device-resident throughput: inputs already in HBM, CUDA events around K steps, max over ranks
  e2e       the same metric through the public class API (`DiscreteHMM.train_batch(host_array, iterations=1)`):
            pinned HOST observations in, H2D + validation + iteration + D2H of the re-estimated parameters per step
  roofline  dominant kernel against its BINDING roof (c4: HBM traffic of the forward stash; secondary roofs beside it)
  cpu_baseline  the oracle (NumPy port of the reference algorithm) timed on the host cores on a bounded sample
  also      driver-timed sub-records of the other BASELINE.json configurations (c2 forward-backward N=8, c3 Viterbi
            N=64, c5 tensor-core forward-backward N=1024), each with ms_per_step / value / roofline / e2e, measured in
            the same process right after the main workload (weak scaling except c5, whose fixed batch is split)
`--impl reference` times the CPU implementation alone (oracle port, all host cores) on the same config.
"""
from __future__ import annotations

import argparse
import ctypes
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
    "c3": ("DiscreteHMM N=64 M=256 Viterbi, B=65536 T=4096 per GPU", 64, 256, 65536, 4096),
    "c4": ("DiscreteHMM N=128 M=1024 Baum-Welch iteration (E-step + NCCL count all-reduce + M-step), B=131072 T=512 per GPU",
           128, 1024, 131072, 512),
    "c5": ("GaussianHMM N=1024 D=1 forward-backward, B=4096 T=1024 (split across ranks)", 1024, 0, 4096, 1024),
}
CFG_INDEX = {"c1": 1, "c2": 2, "c3": 3, "c4": 4, "c5": 5}
METRIC = "state-transitions/sec (B*T*N^2)"


def _json(path):
    try:
        with open(os.path.join(ROOT, path)) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


def peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) for HBM and bf16; the SIMT pipe rates and the tf32
    tensor rate from this repo's own microbenchmarks on the same pool (tools/microbench*.cu -> profiles/)."""
    p = _json("MEASURED_PEAKS.json")
    out = {"hbm_gbs": 6650.0, "bf16_tflops": 1621.0, "bf16_tflops_sustained": 1383.3,
           "hbm_source": "fallback (B200_PROFILING.md)"}
    if p:
        out.update(hbm_gbs=float(p["hbm_gbs"]), bf16_tflops=float(p["bf16_tflops"]),
                   bf16_tflops_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                   hbm_source="measured (MEASURED_PEAKS.json)")
    pipes = _json("profiles/r1_microbench_pipes.json") or {}
    out["ffma2_tflops"] = float(pipes.get("fp32_ffma2_packed_tflops", 67.7))
    out["dadd_gops"] = float(pipes.get("fp64_dadd_gops", 18396.0))
    out["mufu_gops"] = float(pipes.get("mufu_ex2_gops", 4626.7))
    tf = _json("profiles/r2_microbench_tf32.json")
    if tf and tf.get("tf32_tflops"):
        out["tf32_tflops"] = float(tf["tf32_tflops"])
        out["tf32_source"] = "measured tcgen05 kind::tf32 rate (tools/microbench_mma.cu, profiles/r2_microbench_tf32.json)"
    else:
        out["tf32_tflops"] = out["bf16_tflops"] / 2
        out["tf32_source"] = "ASSUMED half of the measured bf16 burst peak (no tf32 microbenchmark result in profiles/)"
    return out


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
CPU_SAMPLE = {"c1": (1, 1000), "c2": (16384, 2048), "c3": (8192, 1024), "c4": (4096, 512), "c5": (32, 32)}


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


# bounded samples for the REAL reference (pure Python, one thread): one sequence, a few seconds each
REAL_REF_T = {"c1": 1000, "c2": 64, "c3": 128, "c4": 32, "c5": 0}


def real_reference_live(kind):
    """The REAL reference timed live on this box's host (one thread, one short sequence) by oracle/time_reference_live.py
    in a child interpreter, so that its `kerehmm` package never shares a process with the product.  None when there is
    nothing to run (no converted copy under oracle/_ref and no /root/reference)."""
    if not REAL_REF_T.get(kind):
        return None
    if not (os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "kerehmm")) or os.path.isdir("/root/reference/kerehmm")):
        return None
    desc, N, M, _, _ = WORKLOADS[kind]
    try:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "time_reference_live.py"), kind, str(N), str(M),
                            str(REAL_REF_T[kind]), str(CFG_INDEX[kind])], capture_output=True, text=True, timeout=120, cwd=ROOT)
        return json.loads(p.stdout.strip().splitlines()[-1])
    except Exception as e:          # the port stands as the baseline; say why the reference itself did not run
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def real_reference_note(kind):
    """The REAL reference (pure Python 2, converted in memory by oracle/ref_loader.py) cannot travel to the GPU box; it
    was timed in the build container on bounded samples (oracle/time_reference.py): one thread, same N/M."""
    d = _json("profiles/r1_reference_cpu_timing.json")
    c = (d or {}).get("cases", {}).get(kind)
    if not c:
        return None
    return {"transitions_per_s": c["transitions_per_s_reference"], "cores": 1, "sample": c["workload"],
            "where": "build container, profiles/r1_reference_cpu_timing.json (the port above is %.0fx faster per core)"
                     % (c["transitions_per_s_oracle_port"] / c["transitions_per_s_reference"])}


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


# ------------------------------------------------------------------------------ workloads
class Workload:
    """One configuration set up on this rank: `step(i)` = one device-resident pass, `e2e_step(i)` = the same through
    the public class API with pinned host buffers, `roofline(ms_per_step, steps)` -> the roofline object."""
    kind = desc = dtype = dominant = l2_note = sharding = scaling = None
    N = M = B = T = 0
    launches_per_step = 0
    h2d = d2h = 0
    host_input = None
    model = obs_h = None

    def release(self):
        pass


def setup_c4(ctx, args):
    import torch
    from kerehmm_b200 import _lib, engine
    from kerehmm_b200 import dist as kdist
    w = Workload()
    w.kind = "c4"
    w.desc, w.N, w.M, w.B, w.T = WORKLOADS["c4"]
    if args.batch:
        w.B = args.batch
    N, M, B, T = w.N, w.M, w.B, w.T
    seed = CFG_INDEX["c4"]
    _, obs_h = make_discrete(N, M, B, T, seed + 100 * ctx.rank)      # this rank's own sequences
    w.model = make_discrete(N, M, 1, 1, seed)[0]                    # the same model on every rank
    w.obs_h = obs_h
    w.host_input = torch.from_numpy(obs_h.view(np.int16)).pin_memory()
    ob = engine.prepare_obs(w.host_input.to(ctx.dev), symbols=True, u16_bits=True, alphabet=M)
    trainer = engine.DiscreteTrainer(w.model._params(), torch.float32, max_iterations=1)
    _lib.call("khmm_estep_tc_timing", 1)
    w.nccl_ranks = None
    if ctx.world > 1:
        kdist.init_comm()
        w.nccl_ranks = kdist.comm_info()[2]

    tc = {"default": True, "bf16x3": "bf16x3", "tf32": "tf32"}[args.precision]
    prec_name = "bf16x3" if args.precision == "default" else args.precision

    def step(i):
        trainer.step(ob, tensor_cores=tc)

    # the class API takes a NumPy uint16 array; this one is a view of the page-locked copy (torch has no first-class uint16)
    host_np = w.host_input.numpy().view(np.uint16)

    def e2e_step(i):
        # public class API: host array in, one Baum-Welch iteration, re-estimated host parameters out
        w.model.train_batch(host_np, iterations=1, tensor_cores=tc)
        return w.model.transitionMatrix

    w.step, w.e2e_step = step, e2e_step
    # libkhmm kernels per step: 2 x bw_round_matrix, bw_forward_tc, bw_backward_tc, bw_fold, mstep_discrete, commit_discrete
    # (+ one torch fill of the 1.2 MB count buffer and, for N > 1, NCCL's all-reduce kernel: not counted)
    w.launches_per_step = 7
    w.h2d, w.d2h = B * T * 2, (N + N * N + N * M) * 8 + 8 + 8
    w.dominant = "bw_backward_tc3_kernel (tcgen05, thread = state, four quarter-tile pipelines per CTA, xi in TMEM, L2 count reductions)"
    w.l2_note = "forward stash of B*T*512 B per step (34 GB at B=131072, >> 126 MB L2): nothing survives between steps"
    w.dtype = ("split bf16 operands (hi + lo, 3 MMAs) / f32 accumulate" if prec_name == "bf16x3" else
               "tf32 multiply / f32 accumulate") + " (float64 counts and M-step)"
    w.sharding = "sequences sharded across ranks, one count all-reduce per iteration (libkhmm communicator over NCCL)"
    w.scaling = "weak"

    def collect(steps):
        """Average device durations of the two E-step kernels over the timed steps (CUDA events inside libkhmm)."""
        n = max(1, min(steps, 32))
        f, b = ctypes.c_float(), ctypes.c_float()
        fs, bs = [], []
        for back in range(n):
            _lib.call("khmm_estep_tc_kernel_ms", back, ctypes.byref(f), ctypes.byref(b))
            fs.append(f.value); bs.append(b.value)
        w.kernel_ms = (float(np.mean(fs)), float(np.mean(bs)))

    w.collect = collect

    def roofline(ms_per_step, steps, pk):
        fwd_ms, bwd_ms = w.kernel_ms
        # algorithmic bytes of the backward kernel: the forward stash read once (512 B per (sequence, step)), two float32
        # scale factors and the 2-byte symbol twice (epilogue + reduction warps); of the forward kernel: stash + scale write, symbol
        bwd_bytes = B * T * (N * 4 + 8 + 4)
        fwd_bytes = B * T * (N * 4 + 4 + 2)
        iter_bytes = fwd_bytes + bwd_bytes
        ach = bwd_bytes / (bwd_ms * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of one backward launch (ncu --set full at B=18944, scaled to
        # this batch; profiles/): traffic ~ algorithmic bytes
        traffic = (5.030011e9 + 0.005003e9) * (B * T) / (18944.0 * 512.0)      # profiles/r2_ncu_c4_bw_backward_tc3_bf16x3_summary.csv
        other, other_ms = w.secondary_result
        return {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": traffic, "kernel": w.dominant, "peak_source": pk["hbm_source"],
                "algorithmic_bytes_per_launch": bwd_bytes, "kernel_ms": bwd_ms,
                "forward_kernel": {"name": "bw_forward_tc_kernel", "kernel_ms": fwd_ms,
                                   "achieved_gbs": fwd_bytes / (fwd_ms * 1e-3) / 1e9,
                                   "frac": fwd_bytes / (fwd_ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
                "iteration": {"algorithmic_bytes": iter_bytes, "hbm_floor_ms": iter_bytes / pk["hbm_gbs"] / 1e6,
                              "ms_per_step": ms_per_step, "frac": iter_bytes / pk["hbm_gbs"] / 1e6 / ms_per_step,
                              "kernels_ms": fwd_ms + bwd_ms,
                              "note": "whole iteration (E-step kernels + fold + all-reduce + M-step + hand-over) "
                                      "against the HBM floor of its fp32 stash write + read"},
                "tensor": {"achieved_tflops": 4.0 * B * T * N * N / (bwd_ms * 1e-3) / 1e12,
                           "peak_tflops": pk["tf32_tflops"], "peak_source": pk["tf32_source"],
                           "note": "secondary roof: MMA1 + MMA2 of the backward kernel counted as 4 flop per transition"},
                "operand_precision": {"this_run": prec_name, "other": other, "other_ms_per_step": other_ms,
                                      "note": "bf16x3 (the API default): re-estimated pi/A/B within 1e-6/2e-6/1e-5 of the float64 "
                                              "oracle at 4,096 sequences; tf32: 6e-5/7e-5/4e-4 (tests/test_gpu_train_loop.py)"},
                "note": "binding roof = HBM (the stash is the only operand that does not fit on chip); the kernel is "
                        "held below it by the L1 / shared-memory data pipe (ncu: LSU wavefronts 87 % of peak -- 4 coalesced "
                        "128-byte accesses per (sequence, 32 states) plus the per-sequence scalars) and runs power-capped "
                        "(clocks.reasons) (DESIGN.md 3.5)"}

    def secondary():
        """The other operand precision, 5 iterations of the same loop (secondary record).  Runs on EVERY rank: the
        iterations contain the count all-reduce."""
        other = "tf32" if prec_name == "bf16x3" else "bf16x3"
        for _ in range(2):
            trainer.step(ob, tensor_cores=other)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            trainer.step(ob, tensor_cores=other)
        e1.record()
        torch.cuda.synchronize()
        w.secondary_result = (other, e0.elapsed_time(e1) / 5)

    w.secondary = secondary
    w.roofline = roofline
    w.trainer = trainer
    return w


def setup_c2(ctx, args):
    import torch
    from kerehmm_b200 import _lib, engine
    w = Workload()
    w.kind = "c2"
    w.desc, w.N, w.M, w.B, w.T = WORKLOADS["c2"]
    if args.batch:
        w.B = args.batch
    N, B, T = w.N, w.B, w.T
    seed = CFG_INDEX["c2"]
    w.model, obs_h = make_gaussian(N, B, T, seed)
    if ctx.rank:      # different sequences per rank, same model
        obs_h = make_gaussian(N, B, T, seed + 100 * ctx.rank)[1]
    w.obs_h = obs_h
    dm = engine.DeviceModel(w.model._params(), torch.float32)
    host_sets = [torch.from_numpy(obs_h).pin_memory(), torch.from_numpy(obs_h[::-1].copy()).pin_memory()]
    dev_sets = [h.to(ctx.dev) for h in host_sets]
    w.host_input = host_sets[0]
    ll = torch.empty((B,), dtype=torch.float64, device=ctx.dev)
    llb = torch.empty((B,), dtype=torch.float64, device=ctx.dev)
    stream = torch.cuda.current_stream()

    def step(i):
        o = dev_sets[i & 1]
        _lib.call("khmm_fwdbwd_loglik_gauss_f32", B, T, N, engine.ptr(dm.pi), engine.ptr(dm.A),
                  engine.ptr(dm.mean), engine.ptr(dm.sd), engine.ptr(o), _lib.F32, None, engine.ptr(ll),
                  engine.ptr(llb), stream.cuda_stream)

    def e2e_step(i):
        return w.model.log_likelihood_batch(host_sets[i & 1])[0]

    w.step, w.e2e_step = step, e2e_step
    w.launches_per_step = 1
    w.h2d, w.d2h = B * T * 4, 2 * B * 8
    w.dominant = "fwdbwd_mma8_kernel<gauss, f32>"
    w.l2_note = "two alternating 134 MB observation sets (268 MB > 126 MB L2)"
    w.dtype = "f32"
    w.sharding = "sequences sharded across ranks, no data-path collective"
    w.scaling = "weak"

    def roofline(ms_per_step, steps, pk):
        algo_bytes = B * T * 4 + 2 * B * 8          # observations once + two float64 results per sequence
        flops = 4.0 * B * T * N * N                 # 2 (forward) + 2 (backward) flop per transition
        ach_tf = flops / (ms_per_step * 1e-3) / 1e12
        ach_gb = algo_bytes / (ms_per_step * 1e-3) / 1e9
        return {"bound": "fp32_simt", "achieved": ach_tf, "peak": pk["ffma2_tflops"], "unit": "TFLOP/s",
                "frac": ach_tf / pk["ffma2_tflops"], "traffic": NCU_TRAFFIC.get("c2"), "kernel": w.dominant,
                "peak_source": "measured packed-FFMA2 rate (tools/microbench.cu, profiles/r1_microbench_pipes.json)",
                "hbm": {"achieved_gbs": ach_gb, "peak_gbs": pk["hbm_gbs"], "frac": ach_gb / pk["hbm_gbs"],
                        "algorithmic_bytes_per_launch": algo_bytes},
                "mufu": {"floor_ms": 16.0 * B * T / (pk["mufu_gops"] * 1e9) * 1e3,
                         "frac": 16.0 * B * T / (pk["mufu_gops"] * 1e9) * 1e3 / ms_per_step},
                "note": "roof = FP32 floor of SURVEY 8(d) (max of the HBM floor and the FP32 floor): achieved counts only "
                        "the 4 N^2 flop per (sequence, step) of the two matrix-vector products.  Since round 2 those "
                        "products run as warp-level mma.sync tiles (split tf32 operands); no pipe is saturated, the "
                        "instruction classes serialise on the scheduler (profiles/r2_c2_mma8_ablation.txt); `mufu` is the "
                        "floor set by the 16 ex2 per (sequence, step) at the measured MUFU rate; `traffic` is the ncu "
                        "figure of the scalar-load instantiation"}

    w.roofline = roofline
    return w


def setup_c3(ctx, args):
    import torch
    from kerehmm_b200 import engine
    w = Workload()
    w.kind = "c3"
    w.desc, w.N, w.M, w.B, w.T = WORKLOADS["c3"]
    if args.batch:
        w.B = args.batch
    N, M, B, T = w.N, w.M, w.B, w.T
    seed = CFG_INDEX["c3"]
    w.model, obs_h = make_discrete(N, M, B, T, seed + 100 * ctx.rank)
    if ctx.rank:
        w.model = make_discrete(N, M, 1, 1, seed)[0]
    w.obs_h = obs_h
    w.host_input = torch.from_numpy(obs_h).pin_memory()
    ob = engine.prepare_obs(w.host_input.to(ctx.dev), symbols=True, alphabet=M)
    p = w.model._params()

    def step(i):
        engine.viterbi(p, ob, torch.uint8)

    def e2e_step(i):
        return w.model.viterbi_batch(w.host_input, path_dtype="uint8")[1]

    w.step, w.e2e_step = step, e2e_step
    w.launches_per_step = 2
    w.h2d, w.d2h = B * T, B * T + B * 8
    w.dominant = "viterbi_screen_kernel<64, uint8> (float32 screen, float64 resolve) + viterbi_backtrace_kernel"
    w.l2_note = "backpointer workspace 17 GB per step (>> L2)"
    w.dtype = "f64"
    w.sharding = "sequences sharded across ranks, no data-path collective"
    w.scaling = "weak"

    def roofline(ms_per_step, steps, pk):
        trans = float(B) * T * N * N
        ach = trans / (ms_per_step * 1e-3)
        fp64_roof = pk["dadd_gops"] * 1e9 / 2.0      # 2 fp64 ops (add + compare/select) per transition
        algo_bytes = B * T * (1 + N + 1 + 1)         # obs + N backpointer bytes written + 1 read + path byte
        ach_gb = algo_bytes / (ms_per_step * 1e-3) / 1e9
        issue_bound = 148 * 4 * 1.965e9 * 32 / 3.15
        return {"bound": "fp64_pipe", "achieved": ach / 1e12, "peak": fp64_roof / 1e12, "unit": "Ttransitions/s",
                "frac": ach / fp64_roof, "traffic": NCU_TRAFFIC.get("c3"), "kernel": w.dominant,
                "peak_source": "SURVEY 8(d) FP64 roof: measured DADD rate / 2 ops per transition "
                               "(profiles/r1_microbench_pipes.json)",
                "hbm": {"achieved_gbs": ach_gb, "peak_gbs": pk["hbm_gbs"], "frac": ach_gb / pk["hbm_gbs"],
                        "algorithmic_bytes_per_launch": algo_bytes},
                "issue_slots": {"issue_bound_transitions_per_s": issue_bound, "frac": ach / issue_bound,
                                "warp_instructions_per_32_transitions": 3.15},
                "note": "bit-exact float64 Viterbi; the float32 screen replaces most FP64-pipe work, so the kernel may "
                        "exceed the plain FP64 roof of the unscreened formulation; it is instruction-issue bound"}

    w.roofline = roofline
    return w


def setup_c5(ctx, args):
    import torch
    from kerehmm_b200 import _lib, engine
    w = Workload()
    w.kind = "c5"
    w.desc, w.N, w.M, w.B, w.T = WORKLOADS["c5"]
    if args.batch:
        w.B = args.batch
    N, T = w.N, w.T
    B = w.B // ctx.world          # STRONG scaling: the configuration fixes the batch at 4096
    w.B_total, w.B = w.B, B
    seed = CFG_INDEX["c5"]
    _, obs_h = make_gaussian(N, B, T, seed + 100 * ctx.rank)
    w.model = make_gaussian(N, 1, 1, seed)[0]
    w.obs_h = obs_h
    w.host_input = torch.from_numpy(obs_h).pin_memory()
    ob = engine.prepare_obs(w.host_input.to(ctx.dev), symbols=False)
    dm = engine.DeviceModel(w.model._params(), torch.float32)
    ll = torch.empty((B,), dtype=torch.float64, device=ctx.dev)
    llb = torch.empty((B,), dtype=torch.float64, device=ctx.dev)
    prec = engine.TC_BF16 if args.operands == "bf16" else engine.TC_TF32
    nbytes = _lib.lib().khmm_fwdbwd_loglik_tc_workspace_bytes(B, N)
    ws = torch.empty((nbytes + 256,), dtype=torch.uint8, device=ctx.dev)
    wsp = ws.data_ptr() + (-ws.data_ptr()) % 256
    stream = torch.cuda.current_stream()

    def step(i):
        _lib.call("khmm_fwdbwd_loglik_tc_gauss_f32", B, T, N, engine.ptr(dm.pi), engine.ptr(dm.A),
                  engine.ptr(dm.AT), engine.ptr(dm.mean), engine.ptr(dm.sd), engine.ptr(ob.tensor), ob.code, wsp,
                  nbytes, engine.ptr(ll), engine.ptr(llb), prec, stream.cuda_stream)

    def e2e_step(i):
        return w.model.log_likelihood_batch(w.host_input, tensor_cores="bf16" if prec else True)[0]

    w.step, w.e2e_step = step, e2e_step
    w.launches_per_step = 5      # 2 x matrix conversion, tc_init, tc_recursion2 (all T-1 steps), tc_final
    w.h2d, w.d2h = B * T * 4, 2 * B * 8
    w.dominant = "tc_recursion2%s_kernel (tcgen05 kind::%s, one cooperative launch for all steps)" % (
        "" if prec else "_pair", "f16 bf16 operands" if prec else "tf32, cta_group::2")
    w.l2_note = "W ping-pong rows + A/AT stay L2 resident by design"
    w.dtype = ("bf16" if prec else "tf32") + " multiply / f32 accumulate"
    w.sharding = "the fixed batch of 4096 sequences is split across ranks, no data-path collective"
    w.scaling = "strong"
    w._ws = ws

    def roofline(ms_per_step, steps, pk):
        tflops = 4.0 * B * T * N * N / (ms_per_step * 1e-3) / 1e12      # per GPU
        bf = bool(prec)
        peak = pk["bf16_tflops"] if bf else pk["tf32_tflops"]
        return {"bound": "tensor", "achieved": tflops, "peak": peak, "unit": "TFLOP/s", "frac": tflops / peak,
                "traffic": NCU_TRAFFIC.get("c5") if bf else None, "kernel": w.dominant,
                "peak_source": "measured bf16 burst peak (MEASURED_PEAKS.json)" if bf else pk["tf32_source"],
                "frac_of_sustained_peak": tflops / pk["bf16_tflops_sustained"] if bf else None,
                "accuracy": "log-likelihood vs float64 oracle at this T: max rel 6.8e-6 (bf16) / 9.2e-7 (tf32), measured"}

    w.roofline = roofline
    return w


def setup_c1(ctx, args):
    import torch
    from kerehmm_b200 import engine
    w = Workload()
    w.kind = "c1"
    w.desc, w.N, w.M, w.B, w.T = WORKLOADS["c1"]
    N, M, B, T = w.N, w.M, w.B, w.T
    w.model, obs_h = make_discrete(N, M, B, T, CFG_INDEX["c1"])
    w.obs_h = obs_h
    p = w.model._params()
    ob = engine.prepare_obs(torch.from_numpy(obs_h.astype(np.int64)).to(ctx.dev), symbols=True, alphabet=M)

    def step(i):          # the three single-sequence calls of the reference's tests, inputs resident on the device
        _, scale, _ = engine.forward(p, ob, torch.float64)
        engine.backward(p, ob, torch.float64, scale)
        engine.viterbi(p, ob)

    def e2e_step(i):      # ... and through the reference-named methods, NumPy in / NumPy out
        o = obs_h[0].astype(np.int64)
        _, scale = w.model.forward(o)
        w.model.backward(o, scale)
        return w.model.viterbi_path(o)[1]

    w.step, w.e2e_step = step, e2e_step
    w.launches_per_step = 4      # forward, backward, Viterbi DP, backtrace
    w.h2d, w.d2h = 3 * B * T * 8, B * T * (2 * N * 8 + 8 + 8)
    w.dominant = "forward/backward/viterbi generic kernels, one CTA: a 1000-step dependency chain (latency, no roofline claim)"
    w.l2_note = "9e3 transitions per step: everything is cache resident; the step is 4 launches of ~1000 dependent steps"
    w.dtype = "f64"
    w.sharding = "single sequence: replicas only"
    w.scaling = "weak"

    def roofline(ms_per_step, steps, pk):
        algo_bytes = B * T * (8 + 3 * N * 8 + N + 8)     # symbols, alpha/scale/beta rows out, backpointers, path
        ach = algo_bytes / (ms_per_step * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": None, "kernel": w.dominant, "peak_source": pk["hbm_source"],
                "note": "latency bound (one sequence): no roofline claim"}

    w.roofline = roofline
    return w


SETUP = {"c1": setup_c1, "c2": setup_c2, "c3": setup_c3, "c4": setup_c4, "c5": setup_c5}

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, ncu --set full at full size (profiles/)
NCU_TRAFFIC = {"c2": 193.742336e6 + 4.065536e6, "c3": 0.309431808e9 + 17.133593e9, "c5": 43.3344e6}


class Ctx:
    pass


def measure(w, ctx, steps, warmup, pk, sample_clocks):
    """Device-resident timing (CUDA events, barrier + synchronize on both sides, max over ranks) and the end-to-end
    timing of one workload -> record dict (rank 0; None elsewhere)."""
    import torch
    import torch.distributed as dist
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(ctx.local) if (sample_clocks and rank == 0) else None
    for i in range(warmup):
        w.step(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    ev0.record(stream)
    for i in range(steps):
        w.step(i)
    ev1.record(stream)
    barrier()
    wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(wall0, wall1) if sampler else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / steps
    transitions = float(w.B) * w.T * w.N * w.N * world
    value = transitions / (ms_per_step * 1e-3)
    if getattr(w, "collect", None) is not None:
        w.collect(steps)
    if getattr(w, "secondary", None) is not None:
        w.secondary()                 # collective-bearing extra measurements: every rank takes part
        barrier()
    roof = w.roofline(ms_per_step, steps, pk) if rank == 0 else None

    # ---- end-to-end through the public API with pinned host buffers
    e2e_steps = max(1, min(steps, 5))
    w.e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        w.e2e_step(i)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = transitions / float(t_e.item())
    # the PCIe floor of the end-to-end number: the same input bytes, pinned host -> device, nothing else
    h2d_copy_ms = h2d_all_ms = None
    if w.host_input is not None:
        dst = torch.empty_like(w.host_input, device=dev)
        best = None
        for _ in range(3):
            barrier()                      # all ranks copy at the same time: the host side is shared
            t0 = time.perf_counter()
            dst.copy_(w.host_input, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        h2d_copy_ms = best * 1e3
        t_c = torch.tensor([best], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_c, op=dist.ReduceOp.MAX)
        h2d_all_ms = float(t_c.item()) * 1e3
        del dst
    if rank != 0:
        return None
    rec = {
        "metric": METRIC, "value": value, "unit": "transitions/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": w.scaling, "vs_baseline": None,
        "dtype": w.dtype, "data": "synthetic (seeded; reference-compatible random model init)",
        "config": {"workload": w.desc, "name": w.kind, "N": w.N, "M": w.M, "B_per_gpu": w.B, "T": w.T,
                   "sharding": w.sharding, "l2": w.l2_note},
        "e2e": {"value": e2e_value, "unit": "transitions/s", "h2d_bytes_per_step": w.h2d, "d2h_bytes_per_step": w.d2h,
                "ms_per_step": float(t_e.item()) * 1e3, "steps": e2e_steps,
                "api": "kerehmm_b200 public class API, pinned host buffers", "h2d_copy_alone_ms": h2d_copy_ms,
                "h2d_copy_alone_gbs": (w.h2d / (h2d_copy_ms * 1e-3) / 1e9) if h2d_copy_ms else None,
                # the host-side ceiling of the end-to-end number: all ranks copying their inputs at once (one NUMA node
                # serves all eight GPUs on this pool's boxes: 232 GB/s aggregate at 8 ranks against 55 GB/s for one)
                "h2d_ceiling_gbs": (world * w.h2d / (h2d_all_ms * 1e-3) / 1e9) if h2d_all_ms else None},
        "gpu_launches": w.launches_per_step * steps,
        "roofline": roof,
    }
    if clocks is not None:
        rec["clocks"] = clocks
    return rec


def multi_gpu_check(ctx):
    """Before timing at N > 1: a small sharded Baum-Welch run (float64 kernels, 2 iterations) through libkhmm's own
    communicator must (a) leave bit-identical parameters on every rank and (b) equal a single-rank run over the whole
    batch to rounding.  The proof travels in the JSON line."""
    import torch
    import torch.distributed as dist
    import kerehmm_b200 as K
    from kerehmm_b200 import dist as kdist
    np.random.seed(77)
    N, M, B, T = 16, 24, 64 * ctx.world + 5, 40
    h = K.DiscreteHMM(N, M, random_transitions=True, random_emissions=True)
    h1 = K.DiscreteHMM(N, M)
    h1.initialProbabilities, h1.transitionMatrix = h.initialProbabilities.copy(), h.transitionMatrix.copy()
    for d, s in zip(h1.emissionDistributions, h.emissionDistributions):
        d.probabilities = s.probabilities.copy()
    obs = np.random.default_rng(78).integers(0, M, size=(B, T))
    lo, hi = kdist.shard_bounds(B)
    hist = h.train_batch(obs[lo:hi], iterations=2, dtype="float64")
    flat = np.concatenate([h.initialProbabilities, h.transitionMatrix.ravel(), h.emission_matrix().ravel(), hist])
    mine = torch.from_numpy(flat).to(ctx.dev)
    allp = [torch.empty_like(mine) for _ in range(ctx.world)]
    dist.all_gather(allp, mine)
    identical = all(bool(torch.equal(allp[0].view(torch.int64), a.view(torch.int64))) for a in allp)
    # single-rank run of the whole batch on this rank, the collective disabled
    hist1 = h1.train_batch(obs, iterations=2, dtype="float64", group=kdist.LOCAL)
    flat1 = np.concatenate([h1.initialProbabilities, h1.transitionMatrix.ravel(), h1.emission_matrix().ravel(), hist1])
    rel = float(np.max(np.abs(flat - flat1) / np.maximum(np.abs(flat1), 1e-300)))
    ok = identical and rel < 1e-9
    info = kdist.comm_info()
    res = {"ranks_bit_identical": identical, "max_rel_vs_single_rank": rel, "passed": ok,
           "nccl_comm_ranks": info[2] if info else None,
           "what": "DiscreteHMM N=16 M=24, %d sequences x T=40 sharded over %d ranks, 2 Baum-Welch iterations, float64 "
                   "kernels, khmm_allreduce_counts" % (B, ctx.world)}
    if not ok:
        raise RuntimeError("multi-GPU parity check failed: %s" % json.dumps(res))
    return res


# ------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from kerehmm_b200 import engine

    ctx = Ctx()
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(ctx.local)
    ctx.dev = torch.device("cuda", ctx.local)
    # the end-to-end leg copies from page-locked host buffers: keep them on the GPU's own NUMA node
    from kerehmm_b200 import dist as _kdist
    ctx.cpus = _kdist.bind_to_device_cpus(ctx.local)
    if ctx.world > 1:
        import datetime
        # a protocol bug must fail the run within minutes, not sit in a collective until the default 10-minute watchdog
        dist.init_process_group("nccl", device_id=ctx.dev, timeout=datetime.timedelta(seconds=180))
    engine.require_cuda()
    pk = peaks()
    check = multi_gpu_check(ctx) if ctx.world > 1 else None

    kind = args.workload
    w = SETUP[kind](ctx, args)
    steps = args.steps
    out = measure(w, ctx, steps, args.warmup, pk, sample_clocks=True)
    if out is not None:
        out["e2e"]["host_cpu_binding"] = ("%d cores local to the GPU (NVML affinity)" % len(ctx.cpus)) if ctx.cpus else "none"
        if check is not None:
            out["multi_gpu_check"] = check
        if kind == "c4":
            if getattr(w, "nccl_ranks", None) is not None:
                out["config"]["nccl_comm_ranks"] = w.nccl_ranks
            # status bits of the device M-steps of all timed iterations (0 = none of the reference's exceptional cases)
            out["config"]["mstep_status"] = int(w.trainer.sticky[0].item())
        if ctx.world == 1 and not args.no_cpu:
            cv, sample, cores, dt = cpu_leg(kind, w.model, w.obs_h, args.cpu_cores)
            out["cpu_baseline"] = {"value": cv, "unit": "transitions/s", "cores": cores, "kind": "port",
                                   "sample": sample, "seconds": dt, "real_reference": real_reference_note(kind),
                                   "real_reference_live": real_reference_live(kind)}
    main_model, main_obs = w.model, w.obs_h
    del w
    engine.release_workspace()
    torch.cuda.empty_cache()

    also = {}
    names = [] if args.also in ("none", "") else [k for k in args.also.split(",") if k and k != kind]
    for k in names:
        if k not in SETUP:
            raise SystemExit("unknown workload in --also: %s" % k)
        wk = SETUP[k](ctx, args)
        # short kernels get more timed steps so that clocks and events see sustained load; long ones fewer
        ksteps = {"c2": max(steps, 50), "c3": max(3, min(steps, 5)), "c5": max(5, min(steps, 10))}.get(k, steps)
        rec = measure(wk, ctx, ksteps, max(3, min(args.warmup, 5)), pk, sample_clocks=False)
        if rec is not None:
            also[k] = {key: rec[key] for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling",
                                                 "dtype", "config", "e2e", "gpu_launches", "roofline")}
        del wk
        engine.release_workspace()
        torch.cuda.empty_cache()
    if out is not None:
        if also:
            out["also"] = also
        print(json.dumps(out), flush=True)
    if ctx.world > 1:
        from kerehmm_b200 import dist as kdist
        kdist.destroy_comm()
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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if kind == "c5":
        B = B // world
    sharding = {"c4": "sequences sharded across ranks, one count all-reduce per iteration (libkhmm communicator over NCCL)",
                "c5": "the fixed batch of 4096 sequences is split across ranks, no data-path collective",
                "c1": "single sequence: replicas only"}.get(kind, "sequences sharded across ranks, no data-path collective")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "transitions/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True,
           "scaling": "strong" if kind == "c5" else "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic (seeded; same generator as the GPU arm)",
           "config": {"workload": desc, "name": kind, "N": N, "M": M, "B_per_gpu": B, "T": T, "sharding": sharding,
                      "l2": "n/a (CPU arm: each step is a bounded sample of the workload, see cpu_baseline.sample)"},
           "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": used, "kind": "port", "sample": sample,
                            "real_reference": real_reference_note(kind), "real_reference_live": real_reference_live(kind)},
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
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--also", default=None,
                    help="comma-separated sub-records measured after the main workload (default: c2,c3,c5 when the "
                         "main workload is c4, else none)")
    ap.add_argument("--batch", type=int, default=0, help="override sequences per GPU (debug)")
    ap.add_argument("--operands", default="bf16", choices=["tf32", "bf16"],
                    help="c5: tensor-core operand precision (bf16: log-likelihoods within 7e-6 of float64 at c5's T=1024, "
                         "measured; tf32: 9e-7)")
    ap.add_argument("--precision", default="default", choices=["default", "bf16x3", "tf32"],
                    help="c4: operand precision of the fused tensor-core E-step (default = the API default, bf16x3)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-cores", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.also is None:
        args.also = "c2,c3,c5" if (args.workload == "c4" and not args.batch) else "none"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
