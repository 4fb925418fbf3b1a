"""CPU oracle for the kereHMM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A float64 NumPy restatement of the reference algorithms (Rabiner-scaled
forward/backward, log-space Viterbi, the reference's non-textbook Baum-Welch
pass and its `smooth_probabilities`).  Every function cites the reference
file:line it follows (paths relative to /root/reference).

Who may import this module: `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` -- as the checker or the
timed CPU baseline only.  The product (`kerehmm_b200`) never imports it and has
no CPU fallback.

Parity status: PINNED.  `oracle/make_golden.py` runs the real reference (py3
token-converted in memory, arithmetic untouched) on seeded inputs and stores
its outputs in `tests/golden/reference_golden.npz`; `tests/test_oracle_golden.py`
checks every function below against those vectors and against the reference's
own known-answer tests (kerehmm/test/discrete_hmm_test.py:48-98).

Third-party arithmetic: Gaussian emission pdfs come from `scipy.stats`
(unpinned by the reference, requirements.txt:1-3; SciPy 1.18.1 here).  The
closed forms below reproduce `norm(loc, scale).pdf` / `multivariate_normal.pdf`
and are checked against SciPy in the tests.
"""
from __future__ import annotations

import numpy as np

DELTA_P = 1.0e-10                       # kerehmm/util.py:4
CONVERGENCE_DELTA_LOG_LIKELIHOOD = 1e-05  # kerehmm/util.py:3


# --------------------------------------------------------------------------- util
def smooth_probabilities(p: np.ndarray) -> np.ndarray:
    """kerehmm/util.py:39-63.  In place; returns its argument.

    1-D: k = #{p < 1e-10}; p -= k*1e-10/(n-k) (ZeroDivisionError when k == n);
    floor at 1e-10; p /= p.sum().
    2-D, per row: normalise; if any entry < 1e-10 the *array of small values*
    z (not its length) is used as the counter (util.py:51-53): p -= z*1e-10/(n-z),
    which broadcasts only for k == 1 or k == n and raises ValueError otherwise;
    floor; normalise again.
    """
    if p.ndim == 2:
        n = p.shape[1]
        for i in range(p.shape[0]):
            row = p[i]
            row /= row.sum()
            small = row < DELTA_P
            if small.any():
                z = row[small]                      # shape (k,)
                row -= z * DELTA_P / (n - z)        # ValueError unless k in {1, n}
                row[row < DELTA_P] = DELTA_P
            row /= row.sum()
    elif p.ndim == 1:
        k = int(np.count_nonzero(p < DELTA_P))
        n = len(p)
        if n - k == 0:
            raise ZeroDivisionError("float division by zero")
        p -= k * DELTA_P / (n - k)
        p[p < DELTA_P] = DELTA_P
        p /= p.sum()
    else:
        raise ValueError("Expected a 1d or 2d array, gotten shape {}.".format(p.shape))
    return p


def random_simplex(size: int, two_d: bool = False) -> np.ndarray:
    """kerehmm/util.py:23-36 (global NumPy RNG, same draw order)."""
    if two_d:
        arr = np.random.uniform(low=DELTA_P, high=.9, size=size ** 2).reshape((size, size))
    else:
        arr = np.random.uniform(low=DELTA_P, high=.9, size=size)
    return smooth_probabilities(arr)


def random_discrete_model(n_states: int, n_symbols: int):
    """Parameters exactly as `DiscreteHMM(N, M, random_transitions=True,
    random_emissions=True)` draws them (abstract_hmm.py:27-28 then
    discrete_hmm.py:19-21 -> distribution.py:31-33): A, pi, then one table per state."""
    A = random_simplex(n_states, two_d=True)
    pi = random_simplex(n_states)
    Bm = np.stack([random_simplex(n_symbols) for _ in range(n_states)])
    return pi, A, Bm


def random_gaussian_model(n_states: int, lower: float = 0.0, upper: float = 100.0):
    """Parameters exactly as `GaussianHMM(N, 1, random_transitions=True,
    random_emissions=True)` (abstract_hmm.py:27-28, gaussian_hmm.py:17-25,
    distribution.py:113-117): means U[lower, upper], "variance" 1 (used as sd)."""
    A = random_simplex(n_states, two_d=True)
    pi = random_simplex(n_states)
    mean = np.array([np.random.uniform([lower], [upper], size=1)[0] for _ in range(n_states)])
    sd = np.ones(n_states)
    return pi, A, mean, sd


# ---------------------------------------------------------------------- emissions
def emissions_discrete(Bm: np.ndarray, obs: np.ndarray) -> np.ndarray:
    """E[..., t, j] = b_j(o_t) = probabilities_j[o_t]  (distribution.py:35-36).
    Bm is (N, M); obs integer (..., T).  Returns (..., T, N)."""
    return np.moveaxis(Bm[:, np.asarray(obs)], 0, -1)


def gaussian_pdf_1d(x, mean, sd):
    """`scipy.stats.norm(loc=mean, scale=variance).pdf(x)` (distribution.py:155):
    the reference's `variance` attribute is passed as the *standard deviation*.
    SciPy: pdf = exp(-z**2/2)/sqrt(2*pi) / scale with z = (x-loc)/scale."""
    z = (np.asarray(x, dtype=np.float64) - mean) / sd
    return np.exp(-z * z / 2.0) / np.sqrt(2.0 * np.pi) / sd


def emissions_gaussian_1d(mean: np.ndarray, sd: np.ndarray, obs: np.ndarray) -> np.ndarray:
    """(..., T) float obs -> (..., T, N)."""
    return gaussian_pdf_1d(np.asarray(obs, dtype=np.float64)[..., None], mean, sd)


def gaussian_pdf_full(x, mean, cov):
    """`multivariate_normal(mean, cov).pdf(x)` (distribution.py:153), full covariance."""
    x = np.asarray(x, dtype=np.float64)
    d = mean.shape[0]
    L = np.linalg.cholesky(cov)
    y = np.linalg.solve(L, (x - mean).reshape(-1, d).T)       # (d, n)
    maha = (y * y).sum(axis=0)
    logdet = 2.0 * np.log(np.diag(L)).sum()
    out = np.exp(-0.5 * (maha + logdet + d * np.log(2.0 * np.pi)))
    return out.reshape(x.shape[:-1])


def emissions_gaussian_full(means: np.ndarray, covs: np.ndarray, obs: np.ndarray) -> np.ndarray:
    """means (N, D), covs (N, D, D), obs (..., T, D) -> (..., T, N)."""
    return np.stack([gaussian_pdf_full(obs, means[j], covs[j]) for j in range(means.shape[0])], axis=-1)


# ------------------------------------------------------------ forward / backward
def forward(pi: np.ndarray, A: np.ndarray, E: np.ndarray):
    """Scaled forward, abstract_hmm.py:250-285.  E is (T, N) or (B, T, N).
    alpha[0] = pi*b(o_0); c_0 = sum; alpha[0] /= c_0 (:258-260);
    alpha[t,j] = (sum_i alpha[t-1,i] A[i,j]) * b_j(o_t) (:268-270); c_t = sum; normalise (:279-280).
    Returns (alpha, scale) with c_t the *sum* (not its reciprocal)."""
    E = np.asarray(E, dtype=np.float64)
    T = E.shape[-2]
    alpha = np.empty_like(E)
    scale = np.empty(E.shape[:-1])
    a = pi * E[..., 0, :]
    c = a.sum(axis=-1)
    scale[..., 0] = c
    alpha[..., 0, :] = a / c[..., None]
    for t in range(1, T):
        a = (alpha[..., t - 1, :] @ A) * E[..., t, :]
        c = a.sum(axis=-1)
        scale[..., t] = c
        alpha[..., t, :] = a / c[..., None]
    return alpha, scale


def backward(A: np.ndarray, E: np.ndarray, scale: np.ndarray) -> np.ndarray:
    """Scaled backward, abstract_hmm.py:327-351 (ghmm convention):
    beta[T-1] = 1 (:337); beta[t,i] = sum_j A[i,j] b_j(o_{t+1}) beta[t+1,j] / c_{t+1} (:343-349)."""
    E = np.asarray(E, dtype=np.float64)
    T = E.shape[-2]
    beta = np.empty_like(E)
    beta[..., T - 1, :] = 1.0
    for t in range(T - 2, -1, -1):
        v = E[..., t + 1, :] * beta[..., t + 1, :]
        beta[..., t, :] = (v @ A.T) / scale[..., t + 1][..., None]
    return beta


def loglik(scale: np.ndarray) -> np.ndarray:
    """True log P(O) = sum_t log c_t."""
    return np.log(scale).sum(axis=-1)


def forward_probability(scale: np.ndarray):
    """Bug-compatible abstract_hmm.py:241-248: `forward()` returns a tuple, so `[-1]` is the
    scale vector and the method returns -logsumexp(scale), not a log-likelihood."""
    return -np.logaddexp.reduce(scale, axis=-1)


def gamma(alpha: np.ndarray, beta: np.ndarray, scale: np.ndarray) -> np.ndarray:
    """abstract_hmm.py:291-297: gamma[t] = alpha[t]*beta[t]*c_t (true gamma times c_t)."""
    return alpha * beta * scale[..., None]


def zi(A: np.ndarray, E: np.ndarray, alpha: np.ndarray, beta: np.ndarray) -> np.ndarray:
    """abstract_hmm.py:299-325: zi[t,i,j] = alpha[t,i] A[i,j] b_j(o_{t+1}) beta[t+1,j]
    (true xi times c_{t+1}).  Single sequence; materialises (T-1, N, N)."""
    return alpha[:-1, :, None] * A[None] * (E[1:] * beta[1:])[:, None, :]


def xi_sum(A: np.ndarray, E: np.ndarray, alpha: np.ndarray, beta: np.ndarray) -> np.ndarray:
    """sum_t zi[t] without materialising: A o (alpha[:-1]^T (E[1:] o beta[1:])).  (..., N, N)."""
    v = E[..., 1:, :] * beta[..., 1:, :]
    return A * np.einsum("...ti,...tj->...ij", alpha[..., :-1, :], v)


# ------------------------------------------------------------------------ Viterbi
def viterbi(pi: np.ndarray, A: np.ndarray, Bm: np.ndarray, obs: np.ndarray):
    """abstract_hmm.py:203-239, bit-exact restatement (float64, one IEEE op per step in the
    reference's association order).
    delta[0] = log(pi) + log(b(o_0)) (:218-219); transitions = delta[t-1] + log(A[:,j]) (:227);
    delta[t,j] = max(transitions) + log(b_j(o_t)) (:228-229); psi = argmax, first index (:230);
    path[-1] = argmax(delta[-1]) (:235); backtrace (:236-237).
    obs is (T,) or (B, T) integer.  Returns (path int64 like obs, logp)."""
    obs = np.asarray(obs)
    single = obs.ndim == 1
    ob = obs[None] if single else obs
    Bn, T = ob.shape
    N = A.shape[0]
    with np.errstate(divide="ignore"):
        logpi, logA, logB = np.log(pi), np.log(A), np.log(Bm)
    psi = np.zeros((Bn, T, N), dtype=np.int32)
    delta = logpi[None, :] + logB[:, ob[:, 0]].T
    for t in range(1, T):
        tr = delta[:, :, None] + logA[None, :, :]          # (B, i, j)
        psi[:, t, :] = tr.argmax(axis=1)
        delta = tr.max(axis=1) + logB[:, ob[:, t]].T
    path = np.empty((Bn, T), dtype=np.int64)
    path[:, -1] = delta.argmax(axis=1)
    logp = delta.max(axis=1)
    rows = np.arange(Bn)
    for t in range(T - 2, -1, -1):
        path[:, t] = psi[rows, t + 1, path[:, t + 1]]
    if single:
        return path[0], logp[0]
    return path, logp


def log_gaussian_pdf_1d(x, mean, sd):
    """log of distribution.py:151-155 for D = 1 as an explicit float64 expression: z = (x - mean)/sd,
    (-0.5 z) z + (-log sd - log sqrt(2 pi)); x (..., T) -> (..., T, N).  EXTENSION: the reference never takes this
    logarithm for Gaussians (its viterbi_path fails on float observations, abstract_hmm.py:234-237)."""
    x = np.asarray(x, dtype=np.float64)[..., None]
    z = (x - mean) / sd
    logc = -np.log(sd) - np.log(np.sqrt(2.0 * np.pi))
    return -0.5 * z * z + logc


def viterbi_gaussian_1d(pi, A, mean, sd, x):
    """The max-plus DP of abstract_hmm.py:203-239 on 1-D Gaussian log-densities (EXTENSION, parity unpinned: the
    reference raises on float observations).  Same association order as `viterbi`: delta[t-1] + log A, max /
    first-index argmax, + log b.  x is (T,) or (B, T) real.  Returns (path int64, logp)."""
    x = np.asarray(x)
    single = x.ndim == 1
    xb = x[None] if single else x
    Bn, T = xb.shape
    N = A.shape[0]
    with np.errstate(divide="ignore"):
        logpi, logA = np.log(pi), np.log(A)
    logE = log_gaussian_pdf_1d(xb, np.asarray(mean, dtype=np.float64), np.asarray(sd, dtype=np.float64))   # (B, T, N)
    psi = np.zeros((Bn, T, N), dtype=np.int32)
    delta = logpi[None, :] + logE[:, 0, :]
    for t in range(1, T):
        tr = delta[:, :, None] + logA[None, :, :]
        psi[:, t, :] = tr.argmax(axis=1)
        delta = tr.max(axis=1) + logE[:, t, :]
    path = np.empty((Bn, T), dtype=np.int64)
    path[:, -1] = delta.argmax(axis=1)
    logp = delta.max(axis=1)
    rows = np.arange(Bn)
    for t in range(T - 2, -1, -1):
        path[:, t] = psi[rows, t + 1, path[:, t + 1]]
    if single:
        return path[0], logp[0]
    return path, logp


# -------------------------------------------------------------------- Baum-Welch
def sequence_stats_discrete(pi, A, Bm, obs):
    """Sufficient statistics of ONE sequence exactly as `DiscreteHMM.do_pass` forms them
    (discrete_hmm.py:117-159), including the gamma[0] aliasing:

      gamma = alpha*beta*c (abstract_hmm.py:294); pi_part = smooth1d(gamma[0]) IN PLACE, so
      gamma[0] becomes pi_part (discrete_hmm.py:132); emis_num[i,k] = sum_{t:o_t=k} gamma[t,i],
      emis_den[i] = sum_t gamma[t,i] (:137-146); xi_sum[i,j] = sum_{t<T-1} zi[t,i,j],
      gam_sum[i] = sum_{t<T-1} gamma[t,i] (:155-157).
    """
    obs = np.asarray(obs)
    M = Bm.shape[1]
    E = emissions_discrete(Bm, obs)
    alpha, scale = forward(pi, A, E)
    beta = backward(A, E, scale)
    g = gamma(alpha, beta, scale)
    pi_part = smooth_probabilities(g[0])            # view: g[0] is overwritten
    emis_num = np.zeros((A.shape[0], M))
    np.add.at(emis_num.T, obs, g)                   # emis_num[:, o_t] += g[t]
    emis_den = g.sum(axis=0)
    xs = xi_sum(A, E, alpha, beta)
    gam_sum = g[:-1].sum(axis=0)
    return dict(pi=pi_part.copy(), xi=xs, gam=gam_sum, emis_num=emis_num, emis_den=emis_den,
                loglik=loglik(scale), nseq=1.0)


def batch_stats_discrete(pi, A, Bm, obs_batch):
    """Multi-sequence extension (the reference handles one sequence per do_pass):
    the per-sequence reference statistics are summed.  With one sequence this is do_pass."""
    acc = None
    for obs in np.asarray(obs_batch):
        s = sequence_stats_discrete(pi, A, Bm, obs)
        if acc is None:
            acc = s
        else:
            for k in acc:
                acc[k] = acc[k] + s[k]
    return acc


def mstep_discrete(stats):
    """M-step, discrete_hmm.py:132-161 applied to (summed) statistics:
    pi_new = pi_sum / nseq (for one sequence exactly smooth1d(gamma[0]), :132);
    B_new[i] = smooth1d(emis_num[i] / emis_den[i]) (:145-146);
    A_new = smooth2d(rownorm(xi / gam[:,None])) (:155-161)."""
    pi_new = stats["pi"] / stats["nseq"]
    B_new = np.empty_like(stats["emis_num"])
    for i in range(B_new.shape[0]):
        row = stats["emis_num"][i] / stats["emis_den"][i]
        B_new[i] = smooth_probabilities(row)
    A_new = stats["xi"] / stats["gam"][:, None]
    A_new /= np.expand_dims(A_new.sum(axis=1), axis=-1)
    A_new = smooth_probabilities(A_new)
    return pi_new, A_new, B_new


def sanity_check_discrete(pi, A, Bm):
    """abstract_hmm.py:137-147 + discrete_hmm.py:24-27."""
    assert np.isclose(pi.sum(), 1)
    for row in range(A.shape[0]):
        assert np.isclose(A[row].sum(), 1)
    for i in range(Bm.shape[0]):
        assert np.isclose(Bm[i].sum(), 1.)


def do_pass_discrete(pi, A, Bm, obs):
    """One reference Baum-Welch pass on one sequence (discrete_hmm.py:95-191)."""
    new = mstep_discrete(sequence_stats_discrete(pi, A, Bm, obs))
    sanity_check_discrete(*new)
    return new


def do_pass_discrete_batch(pi, A, Bm, obs_batch):
    new = mstep_discrete(batch_stats_discrete(pi, A, Bm, obs_batch))
    sanity_check_discrete(*new)
    return new


def train_discrete(pi, A, Bm, obs, iterations=100, auto_stop=True):
    """abstract_hmm.py:400-441: range(iterations+1) passes; stop when the (bug-compatible)
    forward_probability improves by <= 1e-5.  Returns (pi, A, B, passes_done)."""
    obs = np.asarray(obs)
    done = 0
    for _ in range(iterations + 1):
        old_f = forward_probability(forward(pi, A, emissions_discrete(Bm, obs))[1])
        pi, A, Bm = do_pass_discrete(pi, A, Bm, obs)
        done += 1
        new_f = forward_probability(forward(pi, A, emissions_discrete(Bm, obs))[1])
        if auto_stop and new_f - old_f <= CONVERGENCE_DELTA_LOG_LIKELIHOOD:
            break
    return pi, A, Bm, done


# ------------------------------------------------- left-to-right fixtures (Q16)
def strict_left_to_right(n_states: int):
    """abstract_hmm.py:369-395 -> (pi, A)."""
    A = np.full((n_states, n_states), DELTA_P)
    for s in range(n_states):
        A[s, (s + 1) % n_states] = - DELTA_P * (n_states - 1) + 1.
    pi = np.array([1 - DELTA_P * (n_states - 1)] + [DELTA_P] * (n_states - 1))
    return pi, A


def left_to_right(n_states: int):
    """discrete_hmm.py:196-227 -> (pi, A)."""
    A = np.full((n_states, n_states), DELTA_P)
    for s in range(n_states):
        A[s, s] = - DELTA_P * (n_states - 2) + .5
        A[s, (s + 1) % n_states] = - DELTA_P * (n_states - 2) + .5
    pi = np.array([1 - DELTA_P * (n_states - 1)] + [DELTA_P] * (n_states - 1))
    return pi, A


# ----------------------------------------------------- batched CPU baseline legs
def fwdbwd_loglik_batch(pi, A, E):
    """forward + backward over a batch; returns (loglik, alpha, beta).  Used as the timed
    CPU baseline ("port") by bench.py."""
    alpha, scale = forward(pi, A, E)
    beta = backward(A, E, scale)
    return loglik(scale), alpha, beta


# ----------------------------------------------------------------------------- sampling
def simulate_from_uniforms(pi, A, U, Bm=None, mean=None, sd=None):
    """`AbstractHMM.simulate` (abstract_hmm.py:458-469: per step `transition()` :449-456, then `emit()`
    :444-447) for a batch, with the uniform draws given explicitly: U[b, t] = (state draw, emission draw,
    second Box-Muller draw).  `numpy.random.choice(range(n), p=p)` is restated as NumPy implements it:
    cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, side="right").  Discrete emission
    (distribution.py:38-39) is another `choice`; the Gaussian one (distribution.py:157-161,
    `norm(loc=mean, scale=variance).rvs()`) is mean + sd*z with z by Box-Muller -- the reference draws z
    from NumPy's generator instead, so only the distribution, not the stream, is comparable."""
    def cdf(p):
        c = np.cumsum(np.asarray(p, dtype=np.float64), axis=-1)
        return c / c[..., -1:]
    U = np.asarray(U, dtype=np.float64)
    Bn, T = U.shape[:2]
    cpi, cA = cdf(pi), cdf(A)
    cB = cdf(Bm) if Bm is not None else None
    states = np.zeros((Bn, T), dtype=np.int64)
    out = np.zeros((Bn, T), dtype=np.int64 if Bm is not None else np.float64)
    for b in range(Bn):
        cur = None
        for t in range(T):
            row = cpi if cur is None else cA[cur]
            cur = min(int(np.searchsorted(row, U[b, t, 0], side="right")), len(row) - 1)
            states[b, t] = cur
            if Bm is not None:
                out[b, t] = min(int(np.searchsorted(cB[cur], U[b, t, 1], side="right")), cB.shape[1] - 1)
            else:
                z = np.sqrt(-2.0 * np.log(1.0 - U[b, t, 1])) * np.cos(2.0 * np.pi * U[b, t, 2])
                out[b, t] = mean[cur] + sd[cur] * z
    return states, out


# ------------------------------------------------------ Gaussian Baum-Welch (EXTENSION, no reference)
def batch_stats_gauss(pi, A, mean, sd, obs_batch, lengths=None):
    """Textbook-EM statistics for 1-D Gaussian emissions on the reference's scaled recursions.  EXTENSION:
    `GaussianHMM.do_pass` (gaussian_hmm.py:29-124) raises at HEAD, so nothing here is pinned by
    the reference beyond `forward`/`backward`/`smooth_probabilities`, which are.  gamma = alpha*beta;
    xi_t = alpha_t(i) A_ij b_j(x_{t+1}) beta_{t+1}(j) / c_{t+1}; pi part = smooth1d(gamma_0) per sequence."""
    N = A.shape[0]
    acc = dict(pi=np.zeros(N), xi=np.zeros((N, N)), gam=np.zeros(N), s0=np.zeros(N), s1=np.zeros(N), s2=np.zeros(N),
               loglik=0.0, nseq=0.0)
    obs_batch = np.asarray(obs_batch, dtype=np.float64)
    for b in range(obs_batch.shape[0]):
        x = obs_batch[b] if lengths is None else obs_batch[b, :lengths[b]]
        E = emissions_gaussian_1d(mean, sd, x)
        alpha, c = forward(pi, A, E)
        beta = backward(A, E, c)
        g = alpha * beta
        acc["pi"] += smooth_probabilities(g[0].copy())
        acc["xi"] += A * (alpha[:-1].T @ (E[1:] * beta[1:] / c[1:, None]))
        acc["gam"] += g[:-1].sum(axis=0)
        acc["s0"] += g.sum(axis=0)
        acc["s1"] += g.T @ x
        acc["s2"] += g.T @ (x * x)
        acc["loglik"] += loglik(c)
        acc["nseq"] += 1.0
    return acc


def mstep_gauss(stats, sd_min=1e-6):
    """pi = mean of the per-sequence smooth1d(gamma_0); A = smooth1d of every row of rownorm(xi / gam)
    (util.py:56-60); mean = S1/S0; sd = sqrt(max(S2/S0 - mean^2, sd_min^2))."""
    pi_new = stats["pi"] / stats["nseq"]
    A_new = stats["xi"] / stats["gam"][:, None]
    A_new /= A_new.sum(axis=1, keepdims=True)
    for i in range(A_new.shape[0]):          # the 1-D rule per row: the 2-D rule raises for >= 2 tiny entries
        A_new[i] /= A_new[i].sum()
        smooth_probabilities(A_new[i])
    mean = stats["s1"] / stats["s0"]
    sd = np.sqrt(np.maximum(stats["s2"] / stats["s0"] - mean * mean, sd_min * sd_min))
    return pi_new, A_new, mean, sd
