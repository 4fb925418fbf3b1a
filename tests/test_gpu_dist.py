"""Multi-GPU Baum-Welch: two ranks (NCCL), each with its shard of the sequences, must end up with
the same parameters as one rank that sees the whole batch.  Needs >= 2 GPUs (skipped otherwise;
the world_size-2 logic is covered on CPU by tests/test_dist_gloo.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
import kerehmm_b200 as K
from kerehmm_b200 import dist as kdist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
np.random.seed(5)
h = K.DiscreteHMM(12, 20, random_transitions=True, random_emissions=True)
obs = np.random.default_rng(6).integers(0, 20, size=(101, 48))
lo, hi = kdist.shard_bounds(len(obs))
hist = h.train_batch(obs[lo:hi], iterations=3, dtype="float64")
out = dict(rank=rank, lo=lo, hi=hi, hist=hist, A=h.transitionMatrix.tolist(), B=h.emission_matrix().tolist(),
           pi=h.initialProbabilities.tolist(), comm=list(kdist.comm_info()))   # libkhmm's own NCCL communicator
json.dump(out, open(os.path.join(sys.argv[1], "rank%%d.json" %% rank), "w"))
dist.destroy_process_group()
""" % ROOT


def test_two_rank_baum_welch_matches_single_rank(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import kerehmm_b200 as K
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), str(tmp_path)]
    subprocess.run(cmd, check=True, timeout=600)
    r0 = json.load(open(tmp_path / "rank0.json"))
    r1 = json.load(open(tmp_path / "rank1.json"))
    assert (r0["lo"], r0["hi"], r1["lo"], r1["hi"]) == (0, 51, 51, 101)
    assert r0["comm"] == [0, 2, 2] and r1["comm"] == [1, 2, 2]      # khmm_comm_init_rank: NCCL itself counts 2 ranks
    assert r0["A"] == r1["A"] and r0["B"] == r1["B"] and r0["pi"] == r1["pi"]      # bitwise identical ranks
    assert r0["hist"] == r1["hist"]
    np.random.seed(5)
    h = K.DiscreteHMM(12, 20, random_transitions=True, random_emissions=True)
    obs = np.random.default_rng(6).integers(0, 20, size=(101, 48))
    hist = h.train_batch(obs, iterations=3, dtype="float64")
    np.testing.assert_allclose(r0["hist"], hist, rtol=1e-10)
    np.testing.assert_allclose(np.array(r0["A"]), h.transitionMatrix, rtol=1e-9)
    np.testing.assert_allclose(np.array(r0["B"]), h.emission_matrix(), rtol=1e-9)
