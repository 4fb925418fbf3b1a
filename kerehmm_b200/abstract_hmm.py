"""AbstractHMM -- the reference's model container and algorithm surface
(kerehmm/abstract_hmm.py:8-469) with the inner loops replaced by libkhmm CUDA kernels.

Single-sequence methods keep the reference's names, argument meaning, float64 NumPy returns
and error behaviour; `*_batch` methods are the batched extension the engine exists for
(SURVEY.md 8(b)).  Parameters are plain mutable NumPy attributes, re-read on every call.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np
from numpy.random import choice

from . import _lib, engine
from .util import CONVERGENCE_DELTA_LOG_LIKELIHOOD, DELTA_P, random_simplex

SimulationResult = namedtuple("SimulationResult", ["states", "emissions"])


class AbstractHMM(object):
    """Generic HMM: `nStates`, `transitionMatrix` (N,N) row-stochastic, `initialProbabilities`
    (N,), `emissionDistributions` (one object per state), `stateLabels`, `verbose`."""

    def __init__(self, number_of_states, state_labels=None, verbose=False, random_transitions=False,
                 *args, **kwargs):
        self.current_state = None
        if state_labels is None:
            state_labels = ["State_%d" % i for i in range(number_of_states)]
        assert len(set(state_labels)) == number_of_states
        self.nStates = number_of_states
        if not random_transitions:
            self.transitionMatrix = np.full((self.nStates, self.nStates), 1. / self.nStates)
            self.initialProbabilities = np.full(self.nStates, 1. / self.nStates)
        else:
            # same draw order as the reference: transitions first, then initial probabilities
            self.transitionMatrix = random_simplex(self.nStates, two_d=True)
            self.initialProbabilities = random_simplex(self.nStates)
        self.emissionDistributions = np.array([None for _ in range(self.nStates)], dtype=object)
        self.stateLabels = state_labels
        self.verbose = verbose

    # ------------------------------------------------------------------ packing hooks
    def _params(self) -> engine.ModelParams:
        """Pack the current (mutable) attributes for the device.  Concrete classes override."""
        raise NotImplementedError("AbstractHMM doesn't define emissions, override this for concrete implementations.")

    def _symbols(self) -> bool:
        return False

    def _feature_dims(self) -> int:
        return 0

    def _alphabet(self):
        """Table size the symbols index (discrete models), else None."""
        return None

    def _prep(self, observations, lengths=None, single=False, overlap=None):
        obs = observations
        if single:
            import torch
            if isinstance(obs, torch.Tensor):
                obs = obs[None]
            else:
                obs = np.asarray(obs)[None]
        return engine.prepare_obs(obs, symbols=self._symbols(), lengths=lengths, feature_dims=self._feature_dims(),
                                  alphabet=self._alphabet(), overlap=overlap)

    # ---------------------------------------------------------------- state alignment
    def compute_initprob_mapping_score(self, to_hmm, mapping):
        return np.sum(np.abs(to_hmm.initialProbabilities - self.initialProbabilities[mapping])) / self.nStates

    def compute_transition_mapping_score(self, to_hmm, mapping):
        from_trans = self.transitionMatrix[:, mapping][mapping, :]
        return np.sum(np.abs(to_hmm.transitionMatrix - from_trans)) / (self.nStates ** 2)

    def compute_emission_mapping_score(self, to_hmm, mapping):
        raise NotImplementedError("AbstractHMM doesn't define emissions, override this for concrete implementations.")

    def _emission_pair_costs(self, to_hmm):
        """(N, N) matrix c[k][m] = emission distance of this model's state m mapped onto `to_hmm`'s state k; the
        emission mapping score of a permutation is the sum of c[k][mapping[k]].  Concrete classes override."""
        raise NotImplementedError("AbstractHMM doesn't define emissions, override this for concrete implementations.")

    def mapping_score(self, to_hmm, mapping, emission_total=None):
        """The objective `compute_mapping_to` minimises (abstract_hmm.py:63-94 as its docstrings describe it): initial +
        transition score of the permutation plus its emission score divided by the sum of the emission scores of ALL
        N! permutations -- which has the closed form (N-1)! * sum(c), every (k, m) pair occurring in (N-1)! of them."""
        from math import lgamma
        mapping = list(mapping)
        score = self.compute_initprob_mapping_score(to_hmm, mapping) + self.compute_transition_mapping_score(to_hmm, mapping)
        emis = self.compute_emission_mapping_score(to_hmm, mapping)
        if emission_total is None:
            c = self._emission_pair_costs(to_hmm)
            emission_total = (np.log(c.sum()) + lgamma(self.nStates)) if c.sum() > 0 else None
        if emission_total is not None and emis > 0:
            score += float(np.exp(np.log(emis) - emission_total))
        return score

    def compute_mapping_to(self, to_hmm, exact_max_states=8):
        """Best state permutation from this HMM onto `to_hmm` (abstract_hmm.py:63-94): the permutation minimising
        initial + transition + normalised emission scores.

        The reference enumerates all N! permutations (and is broken at HEAD: its second loop iterates an exhausted
        generator, so the emission scores are never added, and it does not run on Python 3).  Here: up to
        `exact_max_states` states the same exhaustive search over the same objective (exact); above that the initial and
        emission terms -- which are sums of per-state costs -- plus the diagonal of the transition term go to the
        Hungarian algorithm (scipy.optimize.linear_sum_assignment, O(N^3)), followed for N <= 64 by pairwise-swap
        descent on the full objective.  The transition term couples pairs of states (a quadratic assignment problem), so
        beyond `exact_max_states` the result is a local optimum; it recovers a relabelled copy of a model exactly.
        Off the hot path."""
        from math import lgamma
        assert self.nStates == to_hmm.nStates
        n = self.nStates
        c_emis = np.asarray(self._emission_pair_costs(to_hmm), dtype=np.float64)
        total = c_emis.sum()
        log_z = (np.log(total) + lgamma(n)) if total > 0 else None          # log((N-1)! * sum(c))
        if n <= exact_max_states:
            from itertools import permutations
            best, best_score = None, None
            for m in permutations(range(n)):
                sc = self.mapping_score(to_hmm, m, emission_total=log_z)
                if best_score is None or sc < best_score:
                    best, best_score = m, sc
            return best
        from scipy.optimize import linear_sum_assignment
        pi_f, pi_t = np.asarray(self.initialProbabilities), np.asarray(to_hmm.initialProbabilities)
        A_f, A_t = np.asarray(self.transitionMatrix), np.asarray(to_hmm.transitionMatrix)
        cost = np.abs(pi_t[:, None] - pi_f[None, :]) / n
        cost = cost + np.abs(np.diag(A_t)[:, None] - np.diag(A_f)[None, :]) / (n ** 2)
        if log_z is not None:
            with np.errstate(divide="ignore", under="ignore"):
                cost = cost + np.exp(np.log(c_emis) - log_z)
        # tie-break on the emission distances themselves: their normalised weight underflows for large N, yet they are
        # what tells relabelled states apart when the initial probabilities are uniform
        cost = cost + 1e-9 * c_emis / max(total, 1e-300) * n
        rows, cols = linear_sum_assignment(cost)
        mapping = [0] * n
        for k, m in zip(rows, cols):
            mapping[k] = int(m)
        if n <= 64:
            best = self.mapping_score(to_hmm, mapping, emission_total=log_z)
            improved = True
            while improved:
                improved = False
                for a in range(n):
                    for b in range(a + 1, n):
                        mapping[a], mapping[b] = mapping[b], mapping[a]
                        sc = self.mapping_score(to_hmm, mapping, emission_total=log_z)
                        if sc < best - 1e-15:
                            best, improved = sc, True
                        else:
                            mapping[a], mapping[b] = mapping[b], mapping[a]
        return tuple(mapping)

    def rearrange(self, mapping):
        mapping = list(mapping)
        self.initialProbabilities = self.initialProbabilities[mapping]
        self.stateLabels = [self.stateLabels[m] for m in mapping]
        self.transitionMatrix = self.transitionMatrix[mapping, :][:, mapping]
        if isinstance(self.emissionDistributions, np.ndarray):
            self.emissionDistributions = self.emissionDistributions[mapping]
        else:
            self.emissionDistributions = [self.emissionDistributions[m] for m in mapping]

    def rearrange_like(self, hmm):
        self.rearrange(self.compute_mapping_to(hmm))

    def initialize_parameters(self, observations):
        raise NotImplementedError()

    # ------------------------------------------------------------------- housekeeping
    def sanity_check(self, sanitize=False, verbose=None):
        """abstract_hmm.py:128-157: pi and every row of A must sum to 1 (np.isclose);
        `sanitize=True` renormalises the rows of A instead of raising."""
        _tmp_verbose = self.verbose
        if verbose is not None:
            self.verbose = verbose
        try:
            assert np.isclose(self.initialProbabilities.sum(), 1)
            if not self.verbose and np.isclose(np.asarray(self.transitionMatrix).sum(axis=1), 1).all():
                return          # fast path: one vectorised check instead of N Python-level ones
            if self.verbose:
                print("SANITY CHECK: transmat\n{}".format(self.transitionMatrix))
            for row in range(self.nStates):
                sum1 = self.transitionMatrix[row].sum()
                if self.verbose:
                    print("SANITY CHECK #{}: Row sum: {}, Column sum: {}".format(
                        row, sum1, self.transitionMatrix[:, row].sum()))
                if not np.isclose(sum1, 1):
                    print("Problem in matrix:\n{}".format(self.transitionMatrix))
                    print("Offending row: {} (sums up to {})".format(self.transitionMatrix[row], sum1))
                    if not sanitize:
                        raise AssertionError()
                    self.transitionMatrix /= self.transitionMatrix.sum(axis=1)[:, np.newaxis]
                    print("Corrected to: {}".format(np.exp(self.transitionMatrix[row])))
        finally:
            self.verbose = _tmp_verbose

    def transition_probability(self, origin, destination):
        """Linear-scale transition probability (the reference's docstring says "log", its code
        returns the linear value, abstract_hmm.py:159-166)."""
        return self.transitionMatrix[origin, destination]

    def l2s(self, label):
        return self.stateLabels.index(label)

    def s2l(self, state):
        return self.stateLabels[state]

    def initial_probability(self, state):
        return self.initialProbabilities[state]

    def backward_probability(self, observations):
        raise NotImplementedError()

    def do_pass(self, observations):
        raise NotImplementedError()

    # ------------------------------------------------------------------ single sequence
    def viterbi_path(self, observations):
        """Rabiner problem 2 (abstract_hmm.py:203-239).  Returns `(path, log_prob)`; `path` has the
        dtype/shape of `observations` (`np.empty_like`, :234), float64 log-prob.  Bit-exact."""
        obs_arr = np.asarray(observations)
        if obs_arr.ndim != 1:
            raise ValueError("viterbi_path takes one sequence; use viterbi_batch for a batch")
        p = self._params()
        if p.kind != "discrete":
            # the reference indexes psi with float observations here and fails (abstract_hmm.py:234-237);
            # Gaussian decoding exists as the batched extension viterbi_batch only
            raise IndexError("viterbi_path needs integer observations (discrete emissions); "
                             "the reference fails on float observations too (abstract_hmm.py:234-237)")
        ob = self._prep(obs_arr, single=True)
        path, logp = engine.viterbi(p, ob)
        out = np.empty_like(obs_arr)
        out[:] = path[0].cpu().numpy()
        return out, float(logp[0].item())

    def forward(self, observations):
        """Scaled forward (abstract_hmm.py:250-285) -> `(alpha[T,N], scale[T])`, float64; rows of
        alpha sum to 1 and `scale[t]` is the normaliser c_t (the sum, not its reciprocal)."""
        t = engine.require_cuda()
        ob = self._prep(observations, single=True)
        alpha, scale, _ = engine.forward(self._params(), ob, t.float64, want_loglik=False)
        alpha, scale = alpha[0].cpu().numpy(), scale[0].cpu().numpy()
        if self.verbose:
            print("alpha = %s" % alpha)
            print("***********")
        return alpha, scale

    def forward_probability(self, observations):
        """Bug-compatible with abstract_hmm.py:241-248: `forward()` returns a tuple, so the
        reference reduces the *scale vector*: -logsumexp(c).  This is NOT a log-likelihood
        (SURVEY.md Q1); use `log_likelihood` for log P(O)."""
        return -np.logaddexp.reduce(self.forward(observations)[-1])

    def log_likelihood(self, observations):
        """True log P(O | model) = sum_t log c_t of one sequence (float64 kernels)."""
        t = engine.require_cuda()
        ob = self._prep(observations, single=True)
        _, _, ll = engine.forward(self._params(), ob, t.float64, want_alpha=False, want_scale=False)
        return float(ll[0].item())

    def backward(self, observations, scale_coefficients):
        """Scaled backward, ghmm convention (abstract_hmm.py:327-351) -> beta[T,N] float64."""
        t = engine.require_cuda()
        ob = self._prep(observations, single=True)
        scale = t.as_tensor(np.ascontiguousarray(scale_coefficients, dtype=np.float64)[None]).to(engine.device())
        if scale.shape != (1, ob.T):
            raise ValueError("scale_coefficients must have one entry per observation")
        beta, _ = engine.backward(self._params(), ob, t.float64, scale)
        return beta[0].cpu().numpy()

    def gamma(self, alpha, beta, scale):
        """abstract_hmm.py:291-297: gamma[t] = alpha[t]*beta[t]*scale[t] (true gamma times c_t)."""
        t = engine.require_cuda()
        dev = engine.device()
        if np.shape(alpha) != np.shape(beta) or np.ndim(alpha) != 2 or np.shape(scale) != (np.shape(alpha)[0],):
            raise ValueError("gamma needs alpha and beta of shape (T, N) and scale of shape (T,)")
        a = t.as_tensor(np.ascontiguousarray(alpha, dtype=np.float64)).to(dev)
        b = t.as_tensor(np.ascontiguousarray(beta, dtype=np.float64)).to(dev)
        c = t.as_tensor(np.ascontiguousarray(scale, dtype=np.float64)).to(dev)
        return engine.gamma(a, b, c, _lib.GAMMA_REFERENCE).cpu().numpy()

    def zi(self, observations, alpha=None, beta=None, scale=None):
        """Rabiner eq. 37 as the reference computes it (abstract_hmm.py:299-325):
        zi[t,i,j] = alpha[t,i] A[i,j] b_j(o_{t+1}) beta[t+1,j] (true xi times c_{t+1}).
        The reference's guards are inverted (Q7): passing alpha/beta makes it RECOMPUTE them, and
        omitting them fails with TypeError -- both kept."""
        t = engine.require_cuda()
        if alpha is None or beta is None:
            raise TypeError("'NoneType' object is not subscriptable")
        p = self._params()
        ob = self._prep(observations, single=True)
        a, c, _ = engine.forward(p, ob, t.float64, want_loglik=False)
        b, _ = engine.backward(p, ob, t.float64, c)
        E = engine.dense_emissions(p, ob)
        A = t.as_tensor(np.ascontiguousarray(p.A)).to(engine.device())
        return engine.zi_single(A, E[0], a[0], b[0]).cpu().numpy()

    def emission_probability(self, state, observation):
        """b_state(observation), or the vector over all states when `state is None`
        (abstract_hmm.py:353-367)."""
        if state is None:
            return np.array([float(self.emissionDistributions[s][observation]) for s in range(self.nStates)])
        return self.emissionDistributions[state][observation]

    def setup_strict_left_to_right(self, set_emissions=False):
        """abstract_hmm.py:369-395: A = DELTA_P everywhere except s -> s+1 (cyclic); pi peaked on 0."""
        n = self.nStates
        self.transitionMatrix[:] = DELTA_P
        for state in range(n):
            self.transitionMatrix[state, (state + 1) % n] = - DELTA_P * (n - 1) + 1.
        self.initialProbabilities = np.array([1 - DELTA_P * (n - 1)] + [DELTA_P] * (n - 1))
        if set_emissions:
            for i, distribution in enumerate(self.emissionDistributions):
                probs = np.zeros_like(distribution.probabilities)
                probs[i] = 1
                distribution.probabilities = probs

    def train(self, observations, iterations=100, auto_stop=True, verbose=False):
        """Baum-Welch loop of abstract_hmm.py:400-441: `range(iterations + 1)` passes (Q6), each
        `forward_probability` -> `do_pass` -> `forward_probability`; with `auto_stop` it stops when
        that (bug-compatible, Q1) quantity improves by <= 1e-5.  Prints like the reference."""
        i, delta_p, old_f, new_f = 0, None, None, None
        for i in range(iterations + 1):
            print('ITERATION #{}'.format(i))
            old_f = self.forward_probability(observations)
            print('Log likelihood = {}'.format(old_f))
            self.do_pass(observations, verbose)
            new_f = self.forward_probability(observations)
            delta_p = new_f - old_f
            print("Is this an improvement? {}".format("Yes" if delta_p > 0 else "No"))
            print('New log likelihood = {}'.format(new_f))
            print('Delta llk = {}'.format(delta_p))
            if auto_stop and delta_p <= CONVERGENCE_DELTA_LOG_LIKELIHOOD:
                break
        print("Finished training at {} iterations.".format(i))
        print("DELTA FORWARD LOG PROB AFTER TRAINING:", delta_p)
        print("{} --> {}".format(old_f, new_f))

    # ------------------------------------------------------------------------ sampling
    def emit(self):
        return self.emissionDistributions[self.current_state].emit()

    def transition(self):
        if self.current_state is not None:
            self.current_state = choice(range(self.nStates), p=self.transitionMatrix[self.current_state, ])
        else:
            self.current_state = choice(range(self.nStates), p=self.initialProbabilities)

    def simulate(self, iterations=1, reset=True):
        """Ancestral sampling with the global NumPy RNG (abstract_hmm.py:444-469)."""
        emissions, states = [], []
        if reset:
            self.current_state = None
        for _ in range(iterations):
            self.transition()
            states.append(self.current_state)
            emissions.append(self.emit())
        return SimulationResult(states=states, emissions=emissions)

    def simulate_batch(self, n_sequences, iterations, seed=0, uniforms=None, to_host=True):
        """`n_sequences` independent `simulate(iterations)` runs on the device -> `SimulationResult`
        with `states` (B, T) int32 and `emissions` (B, T) (int32 symbols / float64 values).  Same
        procedure as the reference (transition, then emit, per step; numpy.random.choice's CDF search)
        driven by a counter-based Philox4x32-10 stream keyed by `seed`; the reference's global-RNG
        stream is not reproduced, so agreement with `simulate` is distributional.  `uniforms`
        (B, T, 3) overrides the generator (tests)."""
        states, em = engine.simulate(self._params(), int(n_sequences), int(iterations), seed, uniforms)
        if to_host:
            states, em = states.cpu().numpy(), em.cpu().numpy()
        return SimulationResult(states=states, emissions=em)

    # ------------------------------------------------------------------ batched engine
    @staticmethod
    def _real(dtype):
        t = engine.torch()
        if dtype in (None, "float32", "f32", np.float32, t.float32):
            return t.float32
        if dtype in ("float64", "f64", np.float64, t.float64):
            return t.float64
        raise ValueError("dtype must be float32 or float64")

    def forward_batch(self, observations, lengths=None, dtype="float32", return_alpha=True):
        """Batched scaled forward.  observations (B, T) [(B, T, D) for multivariate Gaussians].
        -> `(alpha[B,T,N] | None, scale[B,T], loglik[B] float64)`; host in -> NumPy out, CUDA tensor
        in -> CUDA tensors out."""
        ob = self._prep(observations, lengths)
        alpha, scale, ll = engine.forward(self._params(), ob, self._real(dtype), want_alpha=return_alpha)
        return engine.finish(ob, alpha, scale, ll)

    def backward_batch(self, observations, scale, lengths=None, dtype="float32"):
        t = engine.require_cuda()
        ob = self._prep(observations, lengths)
        real = self._real(dtype)
        sc = scale if isinstance(scale, t.Tensor) else t.as_tensor(np.ascontiguousarray(scale))
        sc = sc.to(engine.device(), dtype=real).contiguous()
        beta, _ = engine.backward(self._params(), ob, real, sc)
        return engine.finish(ob, beta)

    def posterior_batch(self, observations, lengths=None, dtype="float32", reference_gamma=False):
        """Forward-backward: -> `(gamma[B,T,N], loglik[B])`.  gamma is the state posterior
        alpha*beta, or the reference's c_t-weighted `gamma()` when `reference_gamma=True`."""
        ob = self._prep(observations, lengths)
        real, p = self._real(dtype), self._params()
        alpha, scale, ll = engine.forward(p, ob, real)
        kind = _lib.GAMMA_REFERENCE if reference_gamma else _lib.GAMMA_POSTERIOR
        _, gam = engine.backward(p, ob, real, scale, alpha=alpha, want_beta=False, want_gamma=True, gamma_kind=kind)
        return engine.finish(ob, gam, ll)

    def log_likelihood_batch(self, observations, lengths=None, dtype="float32", backward_check=True,
                             tensor_cores=True):
        """log P(O_b | model) for every sequence -> `(loglik, loglik_from_backward | None)`.
        float32, N <= 16: the fused forward+backward register kernel.  float32, N a multiple of 128
        (<= 1024) and `tensor_cores=True`: the tcgen05 (tf32) recursion (ragged batches included), forward and
        backward together in one cooperative launch; `tensor_cores="bf16"` uses bf16 operands (twice the MMA
        rate, looser bound on short sequences -- DESIGN.md 3.4).  Otherwise the generic forward kernel (second
        value None)."""
        real, p = self._real(dtype), self._params()
        t = engine.torch()
        if (real == t.float32 and lengths is None and p.N <= _lib.SMALL_N_MAX and p.kind in ("discrete", "gauss")):
            # large host batch: copies of later row blocks overlap the kernel on earlier ones
            out = engine.fwdbwd_loglik_small_host(p, observations, both=backward_check)
            if out is not None:
                return out
        ob = self._prep(observations, lengths)
        if real == t.float32 and p.N <= _lib.SMALL_N_MAX and p.kind in ("discrete", "gauss"):
            ll, llb = engine.fwdbwd_loglik_small(p, ob, both=backward_check)
            return engine.finish(ob, ll, llb)
        if real == t.float32 and tensor_cores and engine.tc_supported(p, ob):
            prec = engine.TC_BF16 if tensor_cores == "bf16" else engine.TC_TF32
            ll, llb = engine.fwdbwd_loglik_tc(p, ob, both=backward_check, precision=prec)
            return engine.finish(ob, ll, llb)
        _, _, ll = engine.forward(p, ob, real, want_alpha=False, want_scale=False)
        return engine.finish(ob, ll, None)

    def viterbi_batch(self, observations, lengths=None, path_dtype=None):
        """Batched bit-exact Viterbi -> `(path[B,T], log_prob[B])`; entries past a sequence's length
        are 0.  `path_dtype` defaults to int64.  EXTENSION: a GaussianHMM with 1-D emissions decodes real
        observations with log b_j(x) = (-0.5 z) z - log(sd_j) - log sqrt(2 pi) (the reference's own viterbi_path
        fails on float observations); bit-exact against oracle.viterbi_gaussian_1d."""
        t = engine.require_cuda()
        # a large page-locked batch is decoded block by block while it is still arriving, and its paths travel back
        # while the next block is decoded (engine.decoding_blocks)
        ob = self._prep(observations, lengths, overlap=engine.decoding_blocks)
        if path_dtype is not None and not isinstance(path_dtype, t.dtype):
            path_dtype = {"uint8": t.uint8, "int16": t.int16, "int32": t.int32, "int64": t.int64}[
                np.dtype(path_dtype).name]
        path, logp = engine.viterbi(self._params(), ob, path_dtype)
        return engine.finish(ob, path, logp)
