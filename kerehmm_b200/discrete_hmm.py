"""DiscreteHMM -- kerehmm/discrete_hmm.py:11-227 on the B200 engine."""
from __future__ import annotations

import numpy as np

from . import engine
from .abstract_hmm import AbstractHMM
from .distribution import DiscreteDistribution
from .util import DELTA_P


class DiscreteHMM(AbstractHMM):
    """HMM with one categorical emission table per state.

    `DiscreteHMM(number_of_states, alphabet_size, state_labels=None, random_emissions=False,
    verbose=False, random_transitions=False)` -- same constructor, attributes and RNG draw order
    as the reference (discrete_hmm.py:16-22)."""

    def __init__(self, number_of_states, alphabet_size, state_labels=None, random_emissions=False, *args, **kwargs):
        super(DiscreteHMM, self).__init__(number_of_states, state_labels, *args, **kwargs)
        self.alphabetSize = alphabet_size
        self.emissionDistributions = np.array(
            [DiscreteDistribution(n=alphabet_size, randomize=random_emissions) for _ in range(self.nStates)])
        self.sanity_check()

    def sanity_check(self):
        """Takes no arguments, like the reference's override (discrete_hmm.py:24-27)."""
        super(DiscreteHMM, self).sanity_check()
        rows = self._shared_rows()
        if rows is not None:
            sums = rows.sum(axis=1)     # one vectorised sum instead of N Python-level ones (0.25 ms at 128 states)
        else:
            sums = np.array([dist.probabilities.sum() for dist in self.emissionDistributions])
        assert np.isclose(sums, 1.).all()

    def _shared_rows(self):
        """The (N, M) matrix whose rows the distribution objects view (installed by `_assign`), or None once any
        distribution has been given another table."""
        rows = getattr(self, "_emission_rows", None)
        dists = self.emissionDistributions
        if rows is None or len(dists) != rows.shape[0]:
            return None
        for d in dists:
            if getattr(d.probabilities, "base", None) is not rows:
                return None
        return rows

    # -------------------------------------------------------------------------- packing
    def emission_matrix(self):
        """State-major (N, M) table assembled from the per-state distribution objects."""
        rows = self._shared_rows()
        if rows is not None:
            return rows.copy()
        return np.stack([np.asarray(d.probabilities, dtype=np.float64) for d in self.emissionDistributions])

    def _params(self):
        return engine.ModelParams(kind="discrete", pi=np.asarray(self.initialProbabilities, dtype=np.float64),
                                  A=np.asarray(self.transitionMatrix, dtype=np.float64), B=self.emission_matrix())

    def _symbols(self):
        return True

    def _alphabet(self):
        return int(self.alphabetSize)

    def _assign(self, pi, A, Bm):
        """Install re-estimated parameters the way do_pass does (discrete_hmm.py:168-173):
        new arrays for pi/A, new distribution objects, a sanity_check after each."""
        self.initialProbabilities = pi
        self.sanity_check()
        rows = np.array(Bm, dtype=np.float64)          # one matrix of our own; distribution i views row i
        dists = np.empty(self.nStates, dtype=object)
        for i in range(self.nStates):
            dists[i] = DiscreteDistribution.from_row(rows[i])
        self._emission_rows = rows
        self.emissionDistributions = dists
        self.sanity_check()
        self.transitionMatrix = A
        self.sanity_check()

    # ------------------------------------------------------------------ state alignment
    def compute_emission_mapping_score(self, to_hmm, mapping):
        score = 0.
        for fro, to in zip(self.emissionDistributions[mapping], to_hmm.emissionDistributions):
            assert isinstance(fro, DiscreteDistribution)
            score += fro.b_distance(to)
        return score

    def _emission_pair_costs(self, to_hmm):
        """c[k][m] = Bhattacharyya distance (distribution.py:41-68: -log sqrt(sum p q)) between this model's state m and
        `to_hmm`'s state k, all pairs at once."""
        P, Q = self.emission_matrix(), to_hmm.emission_matrix()
        with np.errstate(divide="ignore"):
            return -np.log(np.sqrt(Q @ P.T))

    def initialize_parameters(self, observations):
        pass

    # ---------------------------------------------------------------------- Baum-Welch
    def do_pass(self, observations, verbose=False):
        """One Baum-Welch pass on ONE sequence with the reference's exact (non-textbook) weighting
        (discrete_hmm.py:95-191): float64 kernels for forward, the
        backward/E-step statistics, the xi contraction and the M-step.  Mutates
        `initialProbabilities`, `transitionMatrix`, `emissionDistributions`."""
        if verbose:
            print("do_pass() called\n  observation: {}\n  initial: {}\n  transition: {}\n  emission: {}".format(
                observations, self.initialProbabilities, self.transitionMatrix, self.emission_matrix()))
        if np.asarray(observations).ndim != 1:
            raise ValueError("do_pass takes one sequence; use do_pass_batch for a batch")
        t = engine.require_cuda()
        p = self._params()
        ob = self._prep(observations, single=True)
        counts = engine.estep_discrete(p, ob, t.float64, tensor_cores=False)
        self._assign(*engine.mstep_discrete(p, counts))
        if verbose:
            print("do_pass() returning\n  initial: {}\n  transition: {}\n  emission: {}".format(
                self.initialProbabilities, self.transitionMatrix, self.emission_matrix()))

    def estep_batch(self, observations, lengths=None, dtype="float32", chunk=None, counts=None,
                    tensor_cores="auto"):
        """Reference-weighted sufficient statistics summed over a batch as an
        `engine.Counts` on the device -- the buffer that is all-reduced across GPUs.
        tensor_cores: True / False / "auto" / "tf32" / "bf16x3".  The fused tcgen05 E-step covers 8 <= N <= 128
        (padded to 128 states), float32, equal-length and ragged batches; "auto" takes it from 2**22 (sequence, step)
        rows.  True and "auto" run the recursion's contraction on split bf16 operands (three MMAs; counts and
        re-estimated parameters agree with the float64 reference to ~1e-5), "tf32" on one tf32 MMA (~15 % faster
        kernels, ~1e-4 systematic error from the rounding of the transition matrix)."""
        ob = self._prep(observations, lengths)
        return engine.estep_discrete(self._params(), ob, self._real(dtype), counts=counts, chunk=chunk,
                                     tensor_cores=tensor_cores)

    def do_pass_batch(self, observations, lengths=None, dtype="float32", chunk=None, group=None,
                      tensor_cores="auto"):
        """One Baum-Welch pass over a batch of sequences (this rank's shard when a process group is
        initialised: the packed counts are all-reduced before the M-step, which every rank then
        performs identically).  Returns the summed log-likelihood of the batch under the OLD
        parameters."""
        return self.train_batch(observations, lengths, 1, dtype, chunk, group, None, tensor_cores)[0]

    def train_batch(self, observations, lengths=None, iterations=10, dtype="float32", chunk=None, group=None,
                    tol=None, tensor_cores="auto"):
        """`iterations` batched Baum-Welch passes; returns the list of per-pass log-likelihoods.
        Stops early when `tol` is given and the log-likelihood gain drops below it.

        The loop is device-resident (`engine.DiscreteTrainer`): the observations are copied to the GPU once, and
        E-step, count all-reduce, M-step and the parameter hand-over of every iteration are queued without a host
        synchronisation (with `tol` the host reads one double per iteration).  The model attributes are replaced
        once, after the last iteration, the way do_pass installs them (discrete_hmm.py:168-173).  If an iteration
        hits one of the reference's exceptional cases (ZeroDivisionError / ValueError of smooth_probabilities, a
        failed sanity_check) the parameters of the last good iteration are installed and that exception is raised."""
        # parameter uploads first: behind the batch they would wait for the whole of it on the copy engine, and the
        # first E-step is meant to start while the batch is still arriving (overlapped upload, engine.prepare_obs)
        trainer = engine.DiscreteTrainer(self._params(), self._real(dtype), max_iterations=iterations)
        ob = self._prep(observations, lengths, overlap=engine.training_blocks)
        for it in range(iterations):
            trainer.step(ob, chunk=chunk, tensor_cores=tensor_cores, group=group)
            if tol is not None and it > 0:
                h = trainer.history[it - 1:it + 1].cpu().numpy()
                if h[1] - h[0] <= tol:
                    break
        pi, A, Bm, history = trainer.finish()
        if trainer.status and trainer.failed_iteration == 0:
            engine._raise_mstep_status(trainer.status)       # nothing was installed on the device either
        self._assign(pi, A, Bm)
        if trainer.status:
            engine._raise_mstep_status(trainer.status)
        return [float(x) for x in history]

    def setup_left_to_right(self, set_emissions=False):
        """discrete_hmm.py:196-227: A[s,s] = A[s,s+1] = 0.5 - (N-2)*DELTA_P (cyclic), rest DELTA_P."""
        n = self.nStates
        self.transitionMatrix[:] = DELTA_P
        for state in range(n):
            self.transitionMatrix[state, state] = - DELTA_P * (n - 2) + .5
            self.transitionMatrix[state, (state + 1) % n] = - DELTA_P * (n - 2) + .5
        self.initialProbabilities = np.array([1 - DELTA_P * (n - 1)] + [DELTA_P] * (n - 1))
        if set_emissions:
            for state in range(n):
                distribution = self.emissionDistributions[state]
                probs = np.full_like(distribution.probabilities, DELTA_P)
                probs[state] = 1 - (n - 1) * DELTA_P
                distribution.probabilities = probs
