"""world_size-2 test (gloo, CPU) of the multi-GPU plumbing: contiguous sequence shards and the
single sum all-reduce of the packed count buffer per Baum-Welch iteration.  Per-shard statistics
come from the oracle here (no GPU); on the GPU box the same code path all-reduces the device
buffer over NCCL (tests/test_gpu_dist.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kerehmm_b200 import dist as kdist
from oracle import hmm_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def pack(stats, N, M):
    return np.concatenate([stats["pi"], (stats["xi"]).ravel(), stats["gam"], stats["emis_num"].T.ravel(),
                           stats["emis_den"], [stats["loglik"], stats["nseq"], 0.0]])


def unpack(buf, N, M):
    o = 0
    out = {}
    for name, size in (("pi", N), ("xi", N * N), ("gam", N), ("emis_num", M * N), ("emis_den", N), ("tail", 3)):
        out[name] = buf[o:o + size]
        o += size
    return dict(pi=out["pi"], xi=out["xi"].reshape(N, N), gam=out["gam"], emis_num=out["emis_num"].reshape(M, N).T,
                emis_den=out["emis_den"], loglik=out["tail"][0], nseq=out["tail"][1])


def _worker(rank, world, port, N, M, B, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        np.random.seed(7)
        pi, A, Bm = O.random_discrete_model(N, M)
        obs = np.random.default_rng(8).integers(0, M, size=(B, T))
        lo, hi = kdist.shard_bounds(B)
        assert kdist.world() == (rank, world)
        stats = O.batch_stats_discrete(pi, A, Bm, obs[lo:hi])
        buf = torch.from_numpy(pack(stats, N, M))
        kdist.all_reduce_counts(buf)
        new = O.mstep_discrete(unpack(buf.numpy(), N, M))
        q.put((rank, lo, hi, buf.numpy().copy(), [x.copy() for x in new]))
    finally:
        dist.destroy_process_group()


def test_sharded_counts_allreduce_world2():
    N, M, B, T = 4, 5, 7, 30
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, M, B, T, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, hi0, buf0, new0), (r1, lo1, hi1, buf1, new1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 4, 4, 7)                  # contiguous, first rank takes the extra
    assert np.array_equal(buf0, buf1)                           # all-reduce result identical on every rank
    for a, b in zip(new0, new1):
        assert np.array_equal(a, b)                             # -> identical M-step without a broadcast
    np.random.seed(7)
    pi, A, Bm = O.random_discrete_model(N, M)
    obs = np.random.default_rng(8).integers(0, M, size=(B, T))
    full = O.mstep_discrete(O.batch_stats_discrete(pi, A, Bm, obs))
    for a, b in zip(new0, full):
        np.testing.assert_allclose(a, b, rtol=1e-12)
    assert unpack(buf0, N, M)["nseq"] == B


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1 << 20):
        for w in (1, 2, 4, 8):
            spans = [kdist.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert kdist.world() == (0, 1) and not kdist.is_initialized()
    buf = torch.ones(3, dtype=torch.float64)
    assert kdist.all_reduce_counts(buf) is buf                  # single process: no-op
