// Generic scaled forward / backward / E-step kernels: any 1 <= N <= 1024, float or double,
// discrete / Gaussian / dense emissions, ragged batches.
//
// Mapping: one CTA owns S sequences at a time (grid-stride over sequence groups); thread j owns
// state j of all S sequences.  The previous row (alpha_{t-1} or b.beta_{t+1}) of the S sequences
// lives in shared memory as [N][S] so one vector LDS broadcasts the S values of state i to the
// whole CTA; A (forward) / A^T (backward) is read [i][j] with j = threadIdx.x, i.e. coalesced
// and bank-conflict free, from shared memory when N*N fits and through L1/L2 otherwise.
// Per step: N*S FMAs per thread, one block reduction of S sums (forward only).
//
// Reference semantics: forward abstract_hmm.py:250-285, backward :327-351, gamma :291-297,
// do_pass statistics discrete_hmm.py:117-159 (see include/khmm.h).
#include "common.cuh"

namespace khmm {

// Sequences per CTA: 4 share every transition-matrix element a thread loads; small batches (fewer groups of four than
// SMs -- a single sequence of the reference's own tests is the extreme) run one sequence per CTA instead: the
// per-step instruction count of a lone warp drops ~4x (its predicated-off slots are gone) and more SMs get a CTA.
constexpr int kSeqPerCta = 4;
static int seq_per_cta(int64_t B) { return (B + kSeqPerCta - 1) / kSeqPerCta < sm_count() ? 1 : kSeqPerCta; }

template <typename R, int S>
struct RowBuf {  // [N][S] rows in shared memory, vector access of the S values of a state
    R *p;
    __device__ __forceinline__ void store(int j, const R (&v)[S]) {
#pragma unroll
        for (int s = 0; s < S; s++) p[j * S + s] = v[s];
    }
    __device__ __forceinline__ void load(int i, R (&v)[S]) const {
        if (sizeof(R) == 4 && S == 4) {
            float4 q = *reinterpret_cast<const float4 *>(p + i * S);
            v[0] = (R)q.x; v[1] = (R)q.y; v[2] = (R)q.z; v[3] = (R)q.w;
        } else if (sizeof(R) == 8 && S == 4) {
            double2 q0 = *reinterpret_cast<const double2 *>(p + i * S);
            double2 q1 = *reinterpret_cast<const double2 *>(p + i * S + 2);
            v[0] = (R)q0.x; v[1] = (R)q0.y; v[2] = (R)q1.x; v[3] = (R)q1.y;
        } else {
#pragma unroll
            for (int s = 0; s < S; s++) v[s] = p[i * S + s];
        }
    }
};

// Block-wide sums of S values; every thread gets the same result (fixed summation order).
template <typename R, int S>
__device__ __forceinline__ void block_sum(R (&v)[S], R *red /* [S][32] */, int nwarps) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < S; s++) {
        R x = warp_sum(v[s]);
        if (lane == 0) red[s * 32 + w] = x;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < S; s++) {
        R x = 0;
        for (int k = 0; k < nwarps; k++) x += red[s * 32 + k];
        v[s] = x;
    }
}

template <typename R>
__device__ __forceinline__ void load_matrix_smem(R *dst, const R *src, int n) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
}

// --------------------------------------------------------------------------- forward
template <typename R, class Emis, bool A_SMEM, int S>
__global__ void __launch_bounds__(1024)
forward_generic_kernel(int64_t B, int64_t T, int N, const R *__restrict__ pi, const R *__restrict__ A, Emis em,
                       const int32_t *__restrict__ lengths, R *__restrict__ alpha, R *__restrict__ scale,
                       double *__restrict__ loglik) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = blockDim.x, nwarps = NT >> 5, j = threadIdx.x;
    R *rows = reinterpret_cast<R *>(smem_raw);          // [2][NT][S]
    R *red = rows + 2 * NT * S;                         // [2][S][32]
    R *As = red + 2 * S * 32;                           // [N][N] when A_SMEM
    const R *Am = A;
    if (A_SMEM) {
        load_matrix_smem(As, A, N * N);
        Am = As;
    }
    const bool active = j < N;
    const R pij = active ? pi[j] : (R)0;
    const int64_t ngroups = (B + S - 1) / S;
    __syncthreads();

    for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int64_t b0 = g * S;
        int len[S];
        int tmax = 0;
#pragma unroll
        for (int s = 0; s < S; s++) {
            int64_t b = b0 + s;
            len[s] = b < B ? (lengths ? lengths[b] : (int)T) : 0;
            tmax = max(tmax, len[s]);
        }
        double ll[S];
#pragma unroll
        for (int s = 0; s < S; s++) ll[s] = 0.0;
        int cur = 0;
        for (int t = 0; t < tmax; t++) {
            R val[S];
            RowBuf<R, S> prev{rows + cur * NT * S}, next{rows + (cur ^ 1) * NT * S};
            if (t == 0) {
#pragma unroll
                for (int s = 0; s < S; s++) val[s] = pij;
            } else {
#pragma unroll
                for (int s = 0; s < S; s++) val[s] = 0;
                if (active) {
#pragma unroll 4
                    for (int i = 0; i < N; i++) {
                        R aij = Am[i * N + j];
                        R a[S];
                        prev.load(i, a);
#pragma unroll
                        for (int s = 0; s < S; s++) val[s] = fma(a[s], aij, val[s]);
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < S; s++) {
                if (active && t < len[s]) {
                    auto o = em.load(b0 + s, t);
                    val[s] *= em.eval(o, j);
                } else {
                    val[s] = 0;
                }
            }
            R c[S];
#pragma unroll
            for (int s = 0; s < S; s++) c[s] = val[s];
            block_sum<R, S>(c, red + (t & 1) * S * 32, nwarps);
#pragma unroll
            for (int s = 0; s < S; s++) {
                if (t < len[s]) {
                    val[s] = val[s] / c[s];
                    int64_t b = b0 + s;
                    if (alpha && active) alpha[(b * T + t) * N + j] = val[s];
                    if (j == 0) {
                        if (scale) scale[b * T + t] = c[s];
                        // with the scale factors going to memory, log P is summed from them afterwards by
                        // loglik_from_scale_kernel: a float64 log per step on thread 0 is on every step's critical path
                        else ll[s] += log((double)c[s]);
                    }
                }
            }
            next.store(j, val);
            __syncthreads();
            cur ^= 1;
        }
        if (j == 0 && loglik && !scale) {
#pragma unroll
            for (int s = 0; s < S; s++)
                if (b0 + s < B) loglik[b0 + s] = ll[s];
        }
    }
}

// log P(O_b) = sum_t log c_t from the stored scale factors: one warp per sequence, lane l adds t = l, l+32, ... in
// order, then a fixed shuffle tree (deterministic; abstract_hmm.py:241-248 sums the same logs)
template <typename R>
__global__ void loglik_from_scale_kernel(int64_t B, int64_t T, const int32_t *__restrict__ lengths, const R *__restrict__ scale,
                                         double *__restrict__ loglik) {
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const int len = lengths ? lengths[b] : (int)T;
    double acc = 0.0;
    for (int t = lane; t < len; t += 32) acc += log((double)scale[b * T + t]);
    acc = warp_sum(acc);
    if (lane == 0) loglik[b] = acc;
}

// E-step: sum over the batch of log P(O_b) = sum_{b,t} log c_t into the count buffer's tail (a float64 log per step on
// thread 0 of the backward kernel was on every step's critical path)
template <typename R>
__global__ void estep_loglik_kernel(int64_t B, int64_t T, const int32_t *__restrict__ lengths, const R *__restrict__ scale,
                                    double *__restrict__ tail) {
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const int len = lengths ? lengths[b] : (int)T;
    double acc = 0.0;
    for (int t = lane; t < len; t += 32) acc += log((double)scale[b * T + t]);
    acc = warp_sum(acc);
    if (lane == 0 && len > 0) atomicAdd(tail, acc);
}

// ------------------------------------------------------------- backward (+gamma, +E-step)
struct EstepOut {
    double *counts;  // packed, see khmm.h
    int M;
};

template <typename R, class Emis, bool A_SMEM, bool ESTEP, int S>
__global__ void __launch_bounds__(1024)
backward_generic_kernel(int64_t B, int64_t T, int N, const R *__restrict__ AT, Emis em,
                        const int32_t *__restrict__ lengths, const R *__restrict__ scale,
                        const R *__restrict__ alpha, R *__restrict__ beta, R *__restrict__ gamma, int gamma_kind,
                        R *__restrict__ V, EstepOut eo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = blockDim.x, nwarps = NT >> 5, i = threadIdx.x;
    R *rows = reinterpret_cast<R *>(smem_raw);  // [2][NT][S]
    R *red = rows + 2 * NT * S;                 // [2][S][32]
    R *As = red + 2 * S * 32;
    const R *Am = AT;
    if (A_SMEM) {
        load_matrix_smem(As, AT, N * N);
        Am = As;
    }
    const bool active = i < N;
    const int64_t ngroups = (B + S - 1) / S;
    // E-step accumulators owned by state i (flushed once per CTA)
    double acc_pi = 0, acc_gam = 0, acc_den = 0, acc_nseq = 0;    // (sum log c_t: estep_loglik_kernel)
    int flag = 0;
    double *c_pi = nullptr, *c_gam = nullptr, *c_num = nullptr, *c_den = nullptr, *c_tail = nullptr;
    if (ESTEP) {
        c_pi = eo.counts;
        c_gam = c_pi + N + (int64_t)N * N;
        c_num = c_gam + N;
        c_den = c_num + (int64_t)eo.M * N;
        c_tail = c_den + N;
    }
    __syncthreads();

    for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int64_t b0 = g * S;
        int len[S];
        int tmax = 0;
#pragma unroll
        for (int s = 0; s < S; s++) {
            int64_t b = b0 + s;
            len[s] = b < B ? (lengths ? lengths[b] : (int)T) : 0;
            tmax = max(tmax, len[s]);
        }
        R bcur[S];  // beta_t(i) of the step just computed
#pragma unroll
        for (int s = 0; s < S; s++) bcur[s] = 0;
        int cur = 0;
        for (int t = tmax - 1; t >= 0; t--) {
            RowBuf<R, S> vrow{rows + cur * NT * S};
            R acc[S];
#pragma unroll
            for (int s = 0; s < S; s++) acc[s] = 0;
            if (t < tmax - 1) {
                // vrow holds v_{t+1}(j) = b_j(o_{t+1}) beta_{t+1}(j) (0 for sequences not yet started)
                if (active) {
#pragma unroll 4
                    for (int jj = 0; jj < N; jj++) {
                        R aij = Am[jj * N + i];
                        R v[S];
                        vrow.load(jj, v);
#pragma unroll
                        for (int s = 0; s < S; s++) acc[s] = fma(v[s], aij, acc[s]);
                    }
                }
            }
            R ct[S], vnew[S];
#pragma unroll
            for (int s = 0; s < S; s++) {
                int64_t b = b0 + s;
                vnew[s] = 0;
                ct[s] = 1;
                if (t < len[s]) {
                    if (t == len[s] - 1) {
                        bcur[s] = 1;
                    } else {
                        bcur[s] = acc[s] / scale[b * T + t + 1];
                    }
                    if (active) {
                        auto o = em.load(b, t);
                        vnew[s] = em.eval(o, i) * bcur[s];
                        if (beta) beta[(b * T + t) * N + i] = bcur[s];
                    }
                    if (ESTEP || (gamma && gamma_kind == KHMM_GAMMA_REFERENCE)) ct[s] = scale[b * T + t];
                }
            }
            if constexpr (!ESTEP) {
                if (gamma && alpha && active) {
#pragma unroll
                    for (int s = 0; s < S; s++)
                        if (t < len[s]) {
                            int64_t o = ((b0 + s) * T + t) * N + i;
                            gamma[o] = alpha[o] * bcur[s] * ct[s];
                        }
                }
            } else {
                // gamma_ref = alpha*beta*c_t (abstract_hmm.py:294)
                R gref[S];
#pragma unroll
                for (int s = 0; s < S; s++) {
                    gref[s] = 0;
                    if (active && t < len[s]) gref[s] = alpha[((b0 + s) * T + t) * N + i] * bcur[s] * ct[s];
                }
                if (t == 0) {
                    // pi part: smooth1d(gamma[0]) in place (discrete_hmm.py:132, util.py:56-60)
                    R cnt[S];
#pragma unroll
                    for (int s = 0; s < S; s++) cnt[s] = (active && len[s] > 0 && gref[s] < (R)1e-10) ? (R)1 : (R)0;
                    block_sum<R, S>(cnt, red, nwarps);
                    R sm[S];
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        sm[s] = 0;
                        if (len[s] > 0) {
                            int k = (int)cnt[s];
                            if (k >= N) {
                                flag = 1;
                            } else if (active) {
                                gref[s] -= (R)((double)k * 1e-10 / (double)(N - k));
                                if (gref[s] < (R)1e-10) gref[s] = (R)1e-10;
                                sm[s] = gref[s];
                            }
                        }
                    }
                    block_sum<R, S>(sm, red + S * 32, nwarps);
#pragma unroll
                    for (int s = 0; s < S; s++)
                        if (active && len[s] > 0) {
                            gref[s] = gref[s] / sm[s];
                            acc_pi += (double)gref[s];
                        }
                }
#pragma unroll
                for (int s = 0; s < S; s++) {
                    if (active && t < len[s]) {
                        auto o = em.load(b0 + s, t);
                        double w = (double)gref[s];
                        atomicAdd(c_num + (int64_t)o.sym * N + i, w);
                        acc_den += w;
                        if (t < len[s] - 1) acc_gam += w;
                        if (V) V[((b0 + s) * T + t) * N + i] = vnew[s];
                    }
                }
            }
            RowBuf<R, S> vnext{rows + (cur ^ 1) * NT * S};
            vnext.store(i, vnew);
            __syncthreads();
            cur ^= 1;
        }
        if (ESTEP && i == 0) {
#pragma unroll
            for (int s = 0; s < S; s++)
                if (len[s] > 0) acc_nseq += 1.0;
        }
    }
    if (ESTEP) {
        if (active) {
            atomicAdd(c_pi + i, acc_pi);
            atomicAdd(c_gam + i, acc_gam);
            atomicAdd(c_den + i, acc_den);
        }
        if (i == 0) {
            atomicAdd(c_tail + 1, acc_nseq);
        }
        if (flag && i == 0) atomicAdd(c_tail + 2, 1.0);
    }
}

// ------------------------------------------------------------------------ launchers
template <typename R>
static size_t generic_smem(int N, bool a_smem, int S = kSeqPerCta) {
    int NT = ((N + 31) / 32) * 32;
    size_t b = sizeof(R) * ((size_t)2 * NT * S + 2 * S * 32);
    if (a_smem) b += sizeof(R) * (size_t)N * N;
    return b;
}

template <typename R>
static bool a_fits_smem(int N) { return generic_smem<R>(N, true) <= 200 * 1024; }

static int grid_for(int64_t B, int ctas_per_sm, int S = kSeqPerCta) {
    int64_t ngroups = (B + S - 1) / S;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    return (int)(ngroups < cap ? (ngroups > 0 ? ngroups : 1) : cap);
}

static int check_common(int64_t B, int64_t T, int N) {
    KHMM_REQUIRE(B >= 0 && T >= 1 && N >= 1, "bad shape B=%lld T=%lld N=%d", (long long)B, (long long)T, N);
    if (N > 1024) {
        set_error("N=%d > 1024 is not supported", N);
        return KHMM_ERR_UNSUPPORTED;
    }
    return KHMM_OK;
}

template <typename R, class Emis>
static int launch_forward(int64_t B, int64_t T, int N, const R *pi, const R *A, Emis em, const int32_t *lengths,
                          R *alpha, R *scale, double *loglik, khmm_stream_t stream) {
    int rc = check_common(B, T, N);
    if (rc) return rc;
    if (B == 0) return KHMM_OK;
    KHMM_REQUIRE(pi && A, "forward: pi/A must not be NULL");
    int NT = ((N + 31) / 32) * 32;
    bool as = a_fits_smem<R>(N);
    const int S = seq_per_cta(B);
    size_t smem = generic_smem<R>(N, as, S);
    // occupancy-aware grid: small CTAs -> several per SM
    int per_sm = NT <= 64 ? 16 : (NT <= 128 ? 8 : (NT <= 256 ? 4 : (NT <= 512 ? 2 : 1)));
    if (as && smem > 48 * 1024) per_sm = (int)max((size_t)1, min((size_t)per_sm, (size_t)(220 * 1024) / smem));
    int grid = grid_for(B, per_sm, S);
    auto kern = S == 1 ? (as ? forward_generic_kernel<R, Emis, true, 1> : forward_generic_kernel<R, Emis, false, 1>)
                       : (as ? forward_generic_kernel<R, Emis, true, kSeqPerCta> : forward_generic_kernel<R, Emis, false, kSeqPerCta>);
    if (smem > 48 * 1024)
        KHMM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NT, smem, (cudaStream_t)stream>>>(B, T, N, pi, A, em, lengths, alpha, scale, loglik);
    KHMM_LAUNCH_CHECK("forward_generic_kernel");
    if (loglik && scale) {
        const int64_t nthreads = B * 32;
        loglik_from_scale_kernel<R><<<(unsigned)((nthreads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(B, T, lengths, scale, loglik);
        KHMM_LAUNCH_CHECK("loglik_from_scale_kernel");
    }
    return KHMM_OK;
}

template <typename R, class Emis, bool ESTEP>
static int launch_backward(int64_t B, int64_t T, int N, const R *AT, Emis em, const int32_t *lengths, const R *scale,
                           const R *alpha, R *beta, R *gamma, int gamma_kind, R *V, EstepOut eo,
                           khmm_stream_t stream) {
    int rc = check_common(B, T, N);
    if (rc) return rc;
    if (B == 0) return KHMM_OK;
    KHMM_REQUIRE(AT && scale, "backward: AT/scale must not be NULL");
    KHMM_REQUIRE(!gamma || alpha, "backward: gamma output needs alpha");
    KHMM_REQUIRE(!ESTEP || (alpha && eo.counts), "estep: needs alpha and counts");
    int NT = ((N + 31) / 32) * 32;
    bool as = a_fits_smem<R>(N);
    const int S = seq_per_cta(B);
    size_t smem = generic_smem<R>(N, as, S);
    int per_sm = NT <= 64 ? 16 : (NT <= 128 ? 8 : (NT <= 256 ? 4 : (NT <= 512 ? 2 : 1)));
    if (as && smem > 48 * 1024) per_sm = (int)max((size_t)1, min((size_t)per_sm, (size_t)(220 * 1024) / smem));
    int grid = grid_for(B, per_sm, S);
    auto kern = S == 1 ? (as ? backward_generic_kernel<R, Emis, true, ESTEP, 1> : backward_generic_kernel<R, Emis, false, ESTEP, 1>)
                       : (as ? backward_generic_kernel<R, Emis, true, ESTEP, kSeqPerCta>
                             : backward_generic_kernel<R, Emis, false, ESTEP, kSeqPerCta>);
    if (smem > 48 * 1024)
        KHMM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NT, smem, (cudaStream_t)stream>>>(B, T, N, AT, em, lengths, scale, alpha, beta, gamma, gamma_kind, V,
                                                   eo);
    KHMM_LAUNCH_CHECK("backward_generic_kernel");
    if (ESTEP) {
        double *tail = eo.counts + N + (int64_t)N * N + N + (int64_t)eo.M * N + N;      // [loglik, nseq, flags]
        estep_loglik_kernel<R><<<(unsigned)((B * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(B, T, lengths, scale, tail);
        KHMM_LAUNCH_CHECK("estep_loglik_kernel");
    }
    return KHMM_OK;
}

static int check_symbols(int64_t B, int obs_dtype, int M, const void *obs, const void *Bsm) {
    if (B == 0) return KHMM_OK;
    KHMM_REQUIRE(obs && Bsm && M >= 1, "discrete: obs/Bsm NULL or M < 1");
    KHMM_REQUIRE(obs_dtype >= KHMM_U8 && obs_dtype <= KHMM_I64, "discrete: obs dtype code %d is not an integer type",
                 obs_dtype);
    return KHMM_OK;
}
static int check_reals(int64_t B, int obs_dtype, const void *obs, const void *mean, const void *sd) {
    if (B == 0) return KHMM_OK;
    KHMM_REQUIRE(obs && mean && sd, "gauss: obs/mean/sd must not be NULL");
    KHMM_REQUIRE(obs_dtype == KHMM_F32 || obs_dtype == KHMM_F64, "gauss: obs dtype code %d is not f32/f64", obs_dtype);
    return KHMM_OK;
}

static int check_gaussfull(int64_t B, int D, int obs_dtype, const void *obs, const void *means, const void *Linv,
                           const void *lognorm) {
    KHMM_REQUIRE(D >= 1 && D <= KHMM_GAUSS_FULL_DMAX, "gaussfull: D = %d features, supported 1 .. %d (use the dense-emission "
                 "entry points beyond that)", D, KHMM_GAUSS_FULL_DMAX);
    if (B == 0) return KHMM_OK;
    KHMM_REQUIRE(obs && means && Linv && lognorm, "gaussfull: obs/means/Linv/lognorm must not be NULL");
    KHMM_REQUIRE(obs_dtype == KHMM_F32 || obs_dtype == KHMM_F64, "gaussfull: obs dtype code %d is not f32/f64", obs_dtype);
    return KHMM_OK;
}

}  // namespace khmm

using namespace khmm;

#define FWD_DISCRETE(R, name)                                                                                         \
    int name(int64_t B, int64_t T, int N, int M, const R *pi, const R *A, const R *Bsm, const void *obs,              \
             int obs_dtype, const int32_t *lengths, R *alpha, R *scale, double *loglik, khmm_stream_t stream) {       \
        int rc = check_symbols(B, obs_dtype, M, obs, Bsm);                                                               \
        if (rc) return rc;                                                                                            \
        DiscreteEmis<R> em{Bsm, obs, obs_dtype, N, M, T};                                                             \
        return launch_forward<R>(B, T, N, pi, A, em, lengths, alpha, scale, loglik, stream);                          \
    }
#define FWD_GAUSS(R, name)                                                                                            \
    int name(int64_t B, int64_t T, int N, const R *pi, const R *A, const R *mean, const R *sd, const void *obs,       \
             int obs_dtype, const int32_t *lengths, R *alpha, R *scale, double *loglik, khmm_stream_t stream) {       \
        int rc = check_reals(B, obs_dtype, obs, mean, sd);                                                               \
        if (rc) return rc;                                                                                            \
        GaussEmis<R> em{mean, sd, obs, obs_dtype, N, T};                                                              \
        return launch_forward<R>(B, T, N, pi, A, em, lengths, alpha, scale, loglik, stream);                          \
    }
#define FWD_GAUSSFULL(R, name)                                                                                        \
    int name(int64_t B, int64_t T, int N, int D, const R *pi, const R *A, const double *means, const double *Linv,    \
             const double *lognorm, const void *obs, int obs_dtype, const int32_t *lengths, R *alpha, R *scale,       \
             double *loglik, khmm_stream_t stream) {                                                                  \
        int rc = check_gaussfull(B, D, obs_dtype, obs, means, Linv, lognorm);                                         \
        if (rc) return rc;                                                                                            \
        GaussFullEmis<R> em{means, Linv, lognorm, obs, obs_dtype, N, D, T};                                           \
        return launch_forward<R>(B, T, N, pi, A, em, lengths, alpha, scale, loglik, stream);                          \
    }
#define BWD_GAUSSFULL(R, name)                                                                                        \
    int name(int64_t B, int64_t T, int N, int D, const R *AT, const double *means, const double *Linv,                \
             const double *lognorm, const void *obs, int obs_dtype, const int32_t *lengths, const R *scale,           \
             const R *alpha, R *beta, R *gamma, int gamma_kind, khmm_stream_t stream) {                               \
        int rc = check_gaussfull(B, D, obs_dtype, obs, means, Linv, lognorm);                                         \
        if (rc) return rc;                                                                                            \
        GaussFullEmis<R> em{means, Linv, lognorm, obs, obs_dtype, N, D, T};                                           \
        return launch_backward<R, GaussFullEmis<R>, false>(B, T, N, AT, em, lengths, scale, alpha, beta, gamma,       \
                                                           gamma_kind, nullptr, EstepOut{nullptr, 0}, stream);        \
    }
#define FWD_DENSE(R, name)                                                                                            \
    int name(int64_t B, int64_t T, int N, const R *pi, const R *A, const R *E, const int32_t *lengths, R *alpha,      \
             R *scale, double *loglik, khmm_stream_t stream) {                                                        \
        KHMM_REQUIRE(E || B == 0, "dense: E must not be NULL");                                                                 \
        DenseEmis<R> em{E, N, T};                                                                                     \
        return launch_forward<R>(B, T, N, pi, A, em, lengths, alpha, scale, loglik, stream);                          \
    }
#define BWD_DISCRETE(R, name)                                                                                         \
    int name(int64_t B, int64_t T, int N, int M, const R *AT, const R *Bsm, const void *obs, int obs_dtype,           \
             const int32_t *lengths, const R *scale, const R *alpha, R *beta, R *gamma, int gamma_kind,               \
             khmm_stream_t stream) {                                                                                  \
        int rc = check_symbols(B, obs_dtype, M, obs, Bsm);                                                               \
        if (rc) return rc;                                                                                            \
        DiscreteEmis<R> em{Bsm, obs, obs_dtype, N, M, T};                                                             \
        return launch_backward<R, DiscreteEmis<R>, false>(B, T, N, AT, em, lengths, scale, alpha, beta, gamma,        \
                                                          gamma_kind, nullptr, EstepOut{nullptr, 0}, stream);         \
    }
#define BWD_GAUSS(R, name)                                                                                            \
    int name(int64_t B, int64_t T, int N, const R *AT, const R *mean, const R *sd, const void *obs, int obs_dtype,    \
             const int32_t *lengths, const R *scale, const R *alpha, R *beta, R *gamma, int gamma_kind,               \
             khmm_stream_t stream) {                                                                                  \
        int rc = check_reals(B, obs_dtype, obs, mean, sd);                                                               \
        if (rc) return rc;                                                                                            \
        GaussEmis<R> em{mean, sd, obs, obs_dtype, N, T};                                                              \
        return launch_backward<R, GaussEmis<R>, false>(B, T, N, AT, em, lengths, scale, alpha, beta, gamma,           \
                                                       gamma_kind, nullptr, EstepOut{nullptr, 0}, stream);            \
    }
#define BWD_DENSE(R, name)                                                                                            \
    int name(int64_t B, int64_t T, int N, const R *AT, const R *E, const int32_t *lengths, const R *scale,            \
             const R *alpha, R *beta, R *gamma, int gamma_kind, khmm_stream_t stream) {                               \
        KHMM_REQUIRE(E || B == 0, "dense: E must not be NULL");                                                                 \
        DenseEmis<R> em{E, N, T};                                                                                     \
        return launch_backward<R, DenseEmis<R>, false>(B, T, N, AT, em, lengths, scale, alpha, beta, gamma,           \
                                                       gamma_kind, nullptr, EstepOut{nullptr, 0}, stream);            \
    }
#define ESTEP_DISCRETE(R, name)                                                                                       \
    int name(int64_t B, int64_t T, int N, int M, const R *AT, const R *Bsm, const void *obs, int obs_dtype,           \
             const int32_t *lengths, const R *alpha, const R *scale, R *V, double *counts, khmm_stream_t stream) {    \
        int rc = check_symbols(B, obs_dtype, M, obs, Bsm);                                                               \
        if (rc) return rc;                                                                                            \
        DiscreteEmis<R> em{Bsm, obs, obs_dtype, N, M, T};                                                             \
        return launch_backward<R, DiscreteEmis<R>, true>(B, T, N, AT, em, lengths, scale, alpha, nullptr, nullptr,    \
                                                         KHMM_GAMMA_REFERENCE, V, EstepOut{counts, M}, stream);       \
    }

extern "C" {
FWD_DISCRETE(float, khmm_forward_discrete_f32)
FWD_DISCRETE(double, khmm_forward_discrete_f64)
FWD_GAUSS(float, khmm_forward_gauss_f32)
FWD_GAUSS(double, khmm_forward_gauss_f64)
FWD_GAUSSFULL(float, khmm_forward_gaussfull_f32)
FWD_GAUSSFULL(double, khmm_forward_gaussfull_f64)
BWD_GAUSSFULL(float, khmm_backward_gaussfull_f32)
BWD_GAUSSFULL(double, khmm_backward_gaussfull_f64)
FWD_DENSE(float, khmm_forward_dense_f32)
FWD_DENSE(double, khmm_forward_dense_f64)
BWD_DISCRETE(float, khmm_backward_discrete_f32)
BWD_DISCRETE(double, khmm_backward_discrete_f64)
BWD_GAUSS(float, khmm_backward_gauss_f32)
BWD_GAUSS(double, khmm_backward_gauss_f64)
BWD_DENSE(float, khmm_backward_dense_f32)
BWD_DENSE(double, khmm_backward_dense_f64)
ESTEP_DISCRETE(float, khmm_estep_discrete_f32)
ESTEP_DISCRETE(double, khmm_estep_discrete_f64)

size_t khmm_counts_len(int N, int M) { return (size_t)N + (size_t)N * N + N + (size_t)M * N + N + 3; }
}
