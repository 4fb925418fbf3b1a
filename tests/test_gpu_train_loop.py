"""GPU tests of the round-2 host/device plumbing around the hot path: the device-resident Baum-Welch loop
(`engine.DiscreteTrainer`, what `DiscreteHMM.train_batch` runs), the range check of live symbols with the reference's
IndexError / negative-index behaviour, and libkhmm's own NCCL communicator entry points (single rank here; two ranks in
tests/test_gpu_dist.py)."""
import ctypes

import numpy as np
import pytest

from oracle import hmm_oracle as O
from test_gpu_parity import K, discrete_model, random_discrete  # noqa: F401  (K is a fixture)

pytestmark = pytest.mark.gpu


def _oracle_train(pi, A, Bm, obs, iterations):
    hist = []
    for _ in range(iterations):
        stats = O.batch_stats_discrete(pi, A, Bm, obs)
        hist.append(stats["loglik"])
        pi, A, Bm = O.mstep_discrete(stats)
    return pi, A, Bm, hist


@pytest.mark.parametrize("N,M,B,T,dtype,rtol", [(6, 8, 40, 64, "float64", 1e-8), (16, 12, 64, 48, "float32", 2e-4),
                                                (128, 24, 200, 32, "float64", 1e-8)])
def test_device_resident_loop_tracks_the_oracle(K, N, M, B, T, dtype, rtol):
    """`train_batch` keeps parameters on the GPU between iterations; after k iterations they must equal the oracle's
    k E/M passes (the float32 bound loosens with the iteration count: errors feed back through the parameters)."""
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=141)
    obs = np.random.default_rng(142).integers(0, M, size=(B, T))
    hist = h.train_batch(obs, iterations=3, dtype=dtype, tensor_cores=False)
    rpi, rA, rB, rhist = _oracle_train(pi, A, Bm, obs, 3)
    np.testing.assert_allclose(hist, rhist, rtol=1e-9 if dtype == "float64" else 1e-5)
    np.testing.assert_allclose(h.initialProbabilities, rpi, rtol=rtol, atol=1e-13)
    np.testing.assert_allclose(h.transitionMatrix, rA, rtol=rtol, atol=1e-13)
    np.testing.assert_allclose(h.emission_matrix(), rB, rtol=rtol, atol=1e-13)
    h.sanity_check()


def test_device_resident_loop_equals_repeated_single_passes(K):
    """k iterations in one device-resident loop == k calls of do_pass_batch (host round trip each time) in float64: the
    hand-over kernel installs exactly what the M-step wrote.  (Equal to rounding, not bit for bit: the statistics are
    summed with float64 atomics whose order varies from launch to launch.)"""
    h1, _ = random_discrete(K, 12, 20, seed=5)
    h2, _ = random_discrete(K, 12, 20, seed=5)
    obs = np.random.default_rng(6).integers(0, 20, size=(101, 48))
    hist1 = h1.train_batch(obs, iterations=3, dtype="float64")
    hist2 = [h2.do_pass_batch(obs, dtype="float64") for _ in range(3)]
    np.testing.assert_allclose(hist1, hist2, rtol=1e-13)
    np.testing.assert_allclose(h1.transitionMatrix, h2.transitionMatrix, rtol=1e-11)
    np.testing.assert_allclose(h1.emission_matrix(), h2.emission_matrix(), rtol=1e-11)
    np.testing.assert_allclose(h1.initialProbabilities, h2.initialProbabilities, rtol=1e-11)


def test_device_resident_loop_tensor_core_path_and_padding(K):
    """The loop on the fused tcgen05 E-step (N = 128, and N = 40 padded to 128 on the device by the hand-over kernel)
    against the float32 SIMT loop, 2 iterations."""
    for N in (128, 40):
        M, B, T = 32, 384, 40
        h1, _ = random_discrete(K, N, M, seed=9)
        h2, _ = random_discrete(K, N, M, seed=9)
        obs = np.random.default_rng(10).integers(0, M, size=(B, T))
        hist1 = h1.train_batch(obs, iterations=2, tensor_cores=True)
        hist2 = h2.train_batch(obs, iterations=2, tensor_cores=False)
        np.testing.assert_allclose(hist1, hist2, rtol=3e-5)
        np.testing.assert_allclose(h1.transitionMatrix, h2.transitionMatrix, rtol=5e-3)
        np.testing.assert_allclose(h1.emission_matrix(), h2.emission_matrix(), rtol=5e-3)
        h1.sanity_check()


def test_loop_reports_the_reference_exceptions_once_at_the_end(K):
    """A model that drives smooth_probabilities' 2-D ValueError case (two tiny entries in a transition row): the
    device loop keeps the last good parameters and raises after the loop, like do_pass raises inside it."""
    N, M = 4, 5
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=3)
    A = A.copy()
    A[1] = [1e-300, 1e-300, 0.5, 0.5]
    h.transitionMatrix = A
    obs = np.random.default_rng(4).integers(0, M, size=(6, 20))
    with pytest.raises(Exception) as ei:
        O.mstep_discrete(O.batch_stats_discrete(pi, A, Bm, obs))
    with pytest.raises(type(ei.value)):
        h.train_batch(obs, iterations=3, dtype="float64")
    assert np.array_equal(h.transitionMatrix, A)          # nothing was installed


def test_tol_stops_early(K):
    h, _ = random_discrete(K, 5, 6, seed=8)
    obs = np.random.default_rng(9).integers(0, 6, size=(30, 40))
    hist = h.train_batch(obs, iterations=50, dtype="float64", tol=1e9)     # any gain is below this tolerance
    assert len(hist) == 2


# ---------------------------------------------------------------------------------- symbols
def test_out_of_range_symbols_raise_like_numpy_indexing(K):
    """distribution.py:35-36 indexes the table with the symbol: >= M or < -M raises IndexError, negatives wrap."""
    h, (pi, A, Bm) = random_discrete(K, 4, 5, seed=0)
    good = np.random.default_rng(1).integers(0, 5, size=(3, 12))
    bad = good.copy()
    bad[1, 7] = 5
    for call in (h.forward_batch, h.viterbi_batch, h.log_likelihood_batch, h.estep_batch):
        with pytest.raises(IndexError):
            call(bad)
    with pytest.raises(IndexError):
        h.forward(bad[1])
    bad[1, 7] = -6
    with pytest.raises(IndexError):
        h.viterbi_batch(bad)
    # negative symbols wrap: -1 is the last symbol
    neg = good.copy()
    neg[neg == 4] = -1
    a1, s1, l1 = h.forward_batch(good, dtype="float64")
    a2, s2, l2 = h.forward_batch(neg, dtype="float64")
    assert np.array_equal(a1, a2) and np.array_equal(l1, l2)
    p1, lp1 = h.viterbi_batch(good)
    p2, lp2 = h.viterbi_batch(neg)
    assert np.array_equal(p1, p2) and np.array_equal(lp1, lp2)
    # padding past a sequence's length may hold anything
    rag = good.copy()
    rag[0, 6:] = 99
    a3, _, l3 = h.forward_batch(rag, lengths=[6, 12, 12], dtype="float64")
    np.testing.assert_array_equal(l3[1:], l1[1:])
    # uint8 symbols with a 256-symbol table can never be out of range (no check, no sync)
    h256, _ = random_discrete(K, 4, 256, seed=2)
    h256.log_likelihood_batch(np.random.default_rng(3).integers(0, 256, size=(4, 9), dtype=np.uint8))


# ------------------------------------------------------------------------------ communicator
def test_comm_entry_points_single_rank(K):
    """khmm_comm_* with world = 1: id creation, init, info, an all-reduce that is the identity, destroy."""
    import torch
    from kerehmm_b200 import _lib
    from kerehmm_b200 import dist as kdist
    L = _lib.lib()
    ver = ctypes.c_int()
    _lib.call("khmm_comm_nccl_version", ctypes.byref(ver))
    assert ver.value >= 20000
    assert kdist._comm is None
    kdist.init_comm(rank=0, world_size=1)
    try:
        assert kdist.world() == (0, 1)
        r, w, n = kdist.comm_info()
        assert (r, w) == (0, 1) and n in (1, -1)
        buf = torch.arange(7, dtype=torch.float64, device="cuda")
        kdist.all_reduce_counts(buf)
        torch.cuda.synchronize()
        assert buf.cpu().tolist() == list(range(7))
    finally:
        kdist.destroy_comm()
    assert kdist._comm is None
    with pytest.raises(_lib.KhmmError):
        _lib.call("khmm_comm_init_rank", ctypes.create_string_buffer(128), 3, 2, 0, ctypes.byref(ctypes.c_void_p()))
    assert L.khmm_comm_destroy(None) == 0


# ------------------------------------------------------- the two backward kernels of the fused E-step
@pytest.mark.parametrize("M,B,T,ragged", [(16, 128, 1, False), (16, 128, 2, False), (16, 130, 8, False), (48, 300, 33, True),
                                          (1024, 1000, 20, False), (24, 700, 64, True), (64, 148 * 128 + 77, 5, True)])
def test_backward_kernels_of_the_fused_estep_agree(K, M, B, T, ragged):
    """bw_backward_tc3_kernel (the default: thread = state, four quarter-tile pipelines per CTA, emission values and count
    reductions straight from registers) against bw_backward_tc_kernel (thread = sequence) and the float64 oracle: same
    statistics on equal-length and ragged batches, partial tiles, more tiles than CTAs, T = 1 and T = 2, in both operand
    precisions."""
    from kerehmm_b200 import _lib, engine
    N = 128
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=N + M + B + T)
    rng = np.random.default_rng(B * T)
    obs = rng.integers(0, M, size=(B, T))
    lengths = rng.integers(1, T + 1, size=B) if ragged else None
    ob = engine.prepare_obs(obs, symbols=True, lengths=lengths, alphabet=M)
    p = h._params()
    res = {}
    try:
        for version, prec in ((1, engine.TC_TF32), (3, engine.TC_TF32), (1, engine.TC_BF16X3), (3, engine.TC_BF16X3)):
            _lib.call("khmm_debug_bw_backward_version", version)
            res[(version, prec)] = engine.estep_discrete_tc(p, ob, precision=prec).host()
    finally:
        _lib.call("khmm_debug_bw_backward_version", 0)
    st = None
    if B <= 2000:
        if lengths is None:
            st = O.batch_stats_discrete(pi, A, Bm, obs)
        else:
            for b in range(B):
                s1 = O.batch_stats_discrete(pi, A, Bm, obs[b:b + 1, :lengths[b]])
                st = s1 if st is None else {k: st[k] + s1[k] for k in st}
    for (version, prec), c in res.items():
        if version == 1:
            continue
        c1 = res[(1, prec)]
        tag = "version %d precision %d " % (version, prec)
        loose = 5e-4 if prec == engine.TC_TF32 else 2e-5          # against the oracle (tf32: the rounding of A is systematic)
        assert c["nseq"] == B and c["flags"] == 0 and np.isclose(c["loglik"], c1["loglik"], rtol=1e-13), tag
        for key in ("pi", "emis_num", "emis_den"):
            np.testing.assert_allclose(c[key], c1[key], rtol=1e-4, atol=1e-6 * np.abs(c1[key]).max(), err_msg=tag + key)
            if st is not None:
                np.testing.assert_allclose(c[key], st[key], rtol=loose, atol=1e-6 * np.abs(st[key]).max(), err_msg=tag + key)
        if T > 1:
            np.testing.assert_allclose(c["gam"], c1["gam"], rtol=1e-4, atol=1e-5 * np.abs(c1["gam"]).max(), err_msg=tag)
            np.testing.assert_allclose(c["xi_raw"], c1["xi_raw"], rtol=1e-3, atol=1e-6 * c1["xi_raw"].max(), err_msg=tag)
            if st is not None:
                np.testing.assert_allclose(c["xi_raw"] * A, st["xi"], rtol=1e-3, atol=1e-6 * st["xi"].max(), err_msg=tag)


# ----------------------------------------------------- operand precision of the fused E-step (config 4 shape)
def _stats_chunk(args):
    pi, A, Bm, obs = args
    return O.batch_stats_discrete(pi, A, Bm, obs)


def test_config4_reestimated_parameters_against_the_oracle(K):
    """BASELINE config 4's model and sequence shape (N = 128, M = 1024, T = 512) on 4,096 sequences: one Baum-Welch pass on
    the fused tensor-core E-step against the float64 oracle's M-step of its own statistics.  Split bf16 operands (three
    MMAs, the default) hold the re-estimated pi / A / B to float32-class bounds; tf32 operands (one MMA) to the looser
    bound stated here -- the rounding of the transition matrix is a fixed perturbation of the model, which averaging over
    more sequences does not remove."""
    import multiprocessing as mp
    import os
    from kerehmm_b200 import engine
    N, M, B, T = 128, 1024, 4096, 512
    h, (pi, A, Bm) = random_discrete(K, N, M, seed=4)
    obs = np.random.default_rng(1004).integers(0, M, size=(B, T))
    procs = max(1, min(16, os.cpu_count() or 1))
    chunks = [(pi, A, Bm, c) for c in np.array_split(obs, procs)]
    with mp.get_context("fork").Pool(procs) as pool:
        parts = pool.map(_stats_chunk, chunks)
    st = {k: sum(p_[k] for p_ in parts) for k in parts[0]}
    rpi, rA, rB = O.mstep_discrete(st)

    def rel(a, b):
        return float(np.max(np.abs(a - b) / np.abs(b)))

    ob = engine.prepare_obs(obs.astype(np.uint16), symbols=True, alphabet=M)
    p = h._params()
    bounds = {"bf16x3": (5e-6, 1e-5, 5e-5, 5e-5), "tf32": (5e-4, 5e-4, 2e-3, 2e-3)}
    for name, prec in (("bf16x3", engine.TC_BF16X3), ("tf32", engine.TC_TF32)):
        c = engine.estep_discrete_tc(p, ob, precision=prec)
        hc = c.host()
        npi, nA, nB = engine.mstep_discrete(p, c)
        b_pi, b_A, b_B, b_cnt = bounds[name]
        assert abs(hc["loglik"] - st["loglik"]) <= 2e-6 * abs(st["loglik"])
        assert rel(npi, rpi) <= b_pi, (name, "pi", rel(npi, rpi))
        assert rel(nA, rA) <= b_A, (name, "A", rel(nA, rA))
        assert rel(nB, rB) <= b_B, (name, "B", rel(nB, rB))
        assert rel(hc["emis_num"], st["emis_num"]) <= b_cnt, (name, "emis_num")
        assert rel(hc["emis_den"], st["emis_den"]) <= b_cnt and rel(hc["gam"], st["gam"]) <= b_cnt, (name, "gam")
    # the public API takes the default precision and names the other
    h2, _ = random_discrete(K, N, M, seed=4)
    h2.do_pass_batch(obs[:512].astype(np.uint16), tensor_cores=True)
    h3, _ = random_discrete(K, N, M, seed=4)
    h3.do_pass_batch(obs[:512].astype(np.uint16), tensor_cores="tf32")
    st5 = O.batch_stats_discrete(pi, A, Bm, obs[:512])
    _, rA5, _ = O.mstep_discrete(st5)
    assert rel(h2.transitionMatrix, rA5) < rel(h3.transitionMatrix, rA5) and rel(h2.transitionMatrix, rA5) <= 2e-5


def test_overlapped_upload_of_a_pinned_batch(K):
    """`train_batch` on a large page-locked batch of unsigned symbols copies it on a side stream in three row blocks and
    runs the E-step block by block (each block waits for its own copy only; the symbol check is deferred to each
    block): same parameters as the device-resident batch, and an out-of-range symbol in the LAST block still raises
    the reference's IndexError and leaves the model untouched."""
    import torch
    from kerehmm_b200 import engine
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    N, M, T = 128, 300, 224
    B = 4 * 128 * sms + 3000
    rng = np.random.default_rng(171)
    host = torch.from_numpy(rng.integers(0, M, size=(B, T)).astype(np.uint16).view(np.int16)).pin_memory()
    assert host.numel() * 2 >= engine.OVERLAP_MIN_BYTES
    ob = engine.prepare_obs(host.numpy().view(np.uint16), symbols=True, alphabet=M, overlap=engine.training_blocks)
    assert ob.ready is not None and [b1 - b0 for b0, b1, _, _ in ob.ready] == [128 * sms, 256 * sms, B - 384 * sms]
    ob.wait_all()
    assert torch.equal(ob.tensor.cpu(), host)
    h1, _ = random_discrete(K, N, M, seed=172)
    h2, _ = random_discrete(K, N, M, seed=172)
    hist1 = h1.train_batch(host.numpy().view(np.uint16), iterations=2)          # overlapped upload
    hist2 = h2.train_batch(host.numpy().view(np.uint16).copy(), iterations=2)   # pageable copy: plain upload
    np.testing.assert_allclose(hist1, hist2, rtol=1e-9)
    np.testing.assert_allclose(h1.initialProbabilities, h2.initialProbabilities, rtol=2e-5, atol=1e-12)
    np.testing.assert_allclose(h1.transitionMatrix, h2.transitionMatrix, rtol=2e-5, atol=1e-12)
    np.testing.assert_allclose(h1.emission_matrix(), h2.emission_matrix(), rtol=2e-5, atol=1e-12)
    # a consumer that does not work block by block (the SIMT E-step) waits for the whole upload
    sub = host[:128 * sms * 4 + 300, :40].contiguous().pin_memory()
    big_enough = engine.OVERLAP_MIN_BYTES
    engine.OVERLAP_MIN_BYTES = 1 << 20
    try:
        h4, _ = random_discrete(K, N, M, seed=172)
        h5, _ = random_discrete(K, N, M, seed=172)
        h4.train_batch(sub.numpy().view(np.uint16), iterations=1, tensor_cores=False)
        h5.train_batch(sub.numpy().view(np.uint16).copy(), iterations=1, tensor_cores=False)
    finally:
        engine.OVERLAP_MIN_BYTES = big_enough
    np.testing.assert_allclose(h4.transitionMatrix, h5.transitionMatrix, rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(h4.emission_matrix(), h5.emission_matrix(), rtol=1e-6, atol=1e-12)
    bad = host.clone().pin_memory()
    bad[B - 7, 11] = M + 5
    h3, _ = random_discrete(K, N, M, seed=172)
    before = (h3.initialProbabilities.copy(), h3.transitionMatrix.copy(), h3.emission_matrix().copy())
    with pytest.raises(IndexError, match="index %d is out of bounds for axis 0 with size %d" % (M + 5, M)):
        h3.train_batch(bad.numpy().view(np.uint16), iterations=1)
    torch.cuda.synchronize()
    assert np.array_equal(h3.initialProbabilities, before[0]) and np.array_equal(h3.transitionMatrix, before[1])
    assert np.array_equal(h3.emission_matrix(), before[2])


def test_overlapped_upload_and_read_back_of_viterbi(K):
    """`viterbi_batch` on a large page-locked batch decodes it block by block while it is still arriving and sends the
    paths of a block back while the next one is decoded: bit-identical to the plain call, IndexError kept."""
    import torch
    from kerehmm_b200 import engine
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    N, M, T = 8, 16, 900
    B = 16 * 16 * sms + 5000
    rng = np.random.default_rng(181)
    host = torch.from_numpy(rng.integers(0, M, size=(B, T)).astype(np.uint8)).pin_memory()
    assert host.numel() >= engine.OVERLAP_MIN_BYTES
    ob = engine.prepare_obs(host, symbols=True, alphabet=M, overlap=engine.decoding_blocks)
    assert ob.ready is not None and [b1 - b0 for b0, b1, _, _ in ob.ready] == [16 * sms, 48 * sms, B - 128 * sms, 48 * sms, 16 * sms]
    h, _ = random_discrete(K, N, M, seed=182)
    path1, logp1 = h.viterbi_batch(host, path_dtype="uint8")                     # overlapped
    path2, logp2 = h.viterbi_batch(host.numpy().copy(), path_dtype="uint8")      # pageable copy: plain upload
    assert isinstance(path1, np.ndarray) and path1.dtype == np.uint8 and path1.shape == (B, T)
    assert np.array_equal(path1, path2) and np.array_equal(logp1, logp2)
    bad = host.clone().pin_memory()
    bad[B - 3, 7] = M
    with pytest.raises(IndexError, match="index %d is out of bounds for axis 0 with size %d" % (M, M)):
        h.viterbi_batch(bad, path_dtype="uint8")
    torch.cuda.synchronize()
