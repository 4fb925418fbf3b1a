// Baum-Welch for 1-D Gaussian emissions -- an EXTENSION: the reference's
// GaussianHMM.do_pass (gaussian_hmm.py:29-124) is stale log-space code that raises at HEAD (Q9), so there are
// no reference semantics to reproduce.  Specified here as textbook EM on the reference's own scaled
// recursions (abstract_hmm.py:250-351):
//     gamma_t = alpha_t . beta_t                      (the posterior; rows sum to 1)
//     xi_t(i,j) = alpha_t(i) A_ij b_j(x_{t+1}) beta_{t+1}(j) / c_{t+1}
//     pi_new   = mean over sequences of smooth1d(gamma_0)         (util.py:56-60, as the discrete path does)
//     A_new    = smooth1d per row of rownorm(sum_t xi_t / sum_{t<len-1} gamma_t)   (util.py:56-60; the 2-D rule
//                of util.py:46-55 raises for rows with several tiny entries, routine for far-apart Gaussians)
//     mean_i   = sum gamma_t(i) x_t / sum gamma_t(i),  sd_i = sqrt(max(sum gamma_t(i) x_t^2 / sum gamma_t(i) - mean_i^2, sd_min^2))
// (`sd` is what the reference stores in GaussianDistribution.variance and hands to scipy as the scale.)
//
// Call order per chunk: khmm_forward_gauss_* (alpha, scale) -> khmm_backward_gauss_* (beta) ->
// khmm_estep_gauss_* (statistics + V[b][t][j] = b_j(x_t) beta_t(j) / c_t) -> khmm_xi_accumulate_* (alpha, V)
// -> all-reduce -> khmm_mstep_gauss_f64.
#include "common.cuh"

namespace khmm {

template <typename R, int S>
__device__ __forceinline__ void block_sum_s(R (&v)[S], R *red, int nwarps) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < S; s++) v[s] = warp_sum(v[s]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; s++) red[w * S + s] = v[s];
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < S; s++) {
        R x = 0;
        for (int k = 0; k < nwarps; k++) x += red[k * S + s];
        v[s] = x;
    }
}

// thread j = state j; the CTA walks sequences b = blockIdx.x, + gridDim.x, ...
template <typename R>
__global__ void __launch_bounds__(1024)
gauss_estep_kernel(int64_t B, int64_t T, int N, const R *__restrict__ mean, const R *__restrict__ sd,
                   const void *__restrict__ obs, int obs_dtype, const int32_t *__restrict__ lengths,
                   const R *__restrict__ alpha, const R *__restrict__ beta, const R *__restrict__ scale,
                   R *__restrict__ V, double *__restrict__ counts) {
    __shared__ R red[64];
    const int j = threadIdx.x;
    const bool active = j < N;
    const int nwarps = blockDim.x >> 5;
    double *c_pi = counts;
    double *c_gam = c_pi + N + (int64_t)N * N;
    double *c_s0 = c_gam + N, *c_s1 = c_s0 + N, *c_s2 = c_s1 + N, *c_tail = c_s2 + N;
    GaussEmis<R> em{mean, sd, obs, obs_dtype, N, T};
    double a_pi = 0, a_gam = 0, a0 = 0, a1 = 0, a2 = 0, a_ll = 0, a_n = 0;
    int flag = 0;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        const int len = lengths ? lengths[b] : (int)T;
        for (int t = 0; t < len; t++) {
            const int64_t o = (b * T + t) * N + j;
            const R c = scale[b * T + t];
            auto ox = em.load(b, t);
            R g = 0, bt = 0;
            if (active) {
                bt = beta[o];
                g = alpha[o] * bt;
                V[o] = em.eval(ox, j) * bt / c;
                const double x = (double)ox.x, gd = (double)g;
                a0 += gd; a1 += gd * x; a2 += gd * x * x;
                if (t < len - 1) a_gam += gd;
            }
            if (t == 0) {   // pi part: smooth1d(gamma_0), util.py:56-60
                R cnt[1] = {(active && g < (R)1e-10) ? (R)1 : (R)0};
                block_sum_s<R, 1>(cnt, red, nwarps);
                const int k = (int)cnt[0];
                R sm[1] = {0};
                if (k >= N) {
                    flag = 1;
                } else if (active) {
                    g -= (R)((double)k * 1e-10 / (double)(N - k));
                    if (g < (R)1e-10) g = (R)1e-10;
                    sm[0] = g;
                }
                block_sum_s<R, 1>(sm, red + 32, nwarps);
                if (active && k < N) a_pi += (double)(g / sm[0]);
            }
            if (j == 0) a_ll += log((double)c);
        }
        if (j == 0 && len > 0) a_n += 1.0;
    }
    if (active) {
        atomicAdd(c_pi + j, a_pi);
        atomicAdd(c_gam + j, a_gam);
        atomicAdd(c_s0 + j, a0);
        atomicAdd(c_s1 + j, a1);
        atomicAdd(c_s2 + j, a2);
    }
    if (j == 0) {
        atomicAdd(c_tail, a_ll);
        atomicAdd(c_tail + 1, a_n);
        if (flag) atomicAdd(c_tail + 2, 1.0);
    }
}

__device__ __forceinline__ double gsum(double v, double *red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double x = 0;
    for (int k = 0; k < nw; k++) x += red[k];
    return x;
}

// blocks [0, N): transition rows (xi*A/gam, row-normalise twice as the discrete M-step does, smooth1d);
// block N: pi, means and standard deviations.
__global__ void __launch_bounds__(256)
mstep_gauss_kernel(int N, const double *__restrict__ counts, const double *__restrict__ A, double sd_min,
                   double *__restrict__ pi_new, double *__restrict__ A_new, double *__restrict__ mean_new,
                   double *__restrict__ sd_new, int32_t *__restrict__ status) {
    __shared__ double red[32];
    const double *c_pi = counts, *c_xi = c_pi + N, *c_gam = c_xi + (int64_t)N * N;
    const double *c_s0 = c_gam + N, *c_s1 = c_s0 + N, *c_s2 = c_s1 + N, *c_tail = c_s2 + N;
    const double DP = 1e-10;
    const int tid = threadIdx.x, nt = blockDim.x, blk = blockIdx.x;
    int st = 0;
    if (blk == N) {
        const double nseq = c_tail[1];
        double s = 0;
        for (int i = tid; i < N; i += nt) {
            double v = c_pi[i] / nseq;
            pi_new[i] = v;
            s += v;
            double m = c_s1[i] / c_s0[i];
            double var = c_s2[i] / c_s0[i] - m * m;
            mean_new[i] = m;
            sd_new[i] = sqrt(var > sd_min * sd_min ? var : sd_min * sd_min);
            if (!(c_s0[i] > 0.0)) st |= 8;   // a state that explains no observation: mean/sd undefined
        }
        s = gsum(s, red);
        if (!(fabs(s - 1.0) <= 1e-8 + 1e-5)) st |= 4;
        if (c_tail[2] != 0.0) st |= 1;
    } else {
        const int i = blk;
        double *row = A_new + (int64_t)i * N;
        const double gam = c_gam[i];
        double s = 0;
        for (int j = tid; j < N; j += nt) {
            double v = c_xi[(int64_t)i * N + j] * A[(int64_t)i * N + j] / gam;
            row[j] = v;
            s += v;
        }
        s = gsum(s, red);
        double s1 = 0;
        for (int j = tid; j < N; j += nt) {
            double v = row[j] / s;
            row[j] = v;
            s1 += v;
        }
        s1 = gsum(s1, red);
        // row-wise smooth1d (util.py:56-60).  The reference's 2-D smoothing raises ValueError as soon as a row
        // has 2 <= k < n entries below 1e-10 (util.py:51-53) -- routine for Gaussian states that are far apart --
        // so this extension applies the well-defined 1-D rule to every row instead.
        double cnt = 0;
        for (int j = tid; j < N; j += nt) {
            double v = row[j] / s1;
            row[j] = v;
            if (v < DP) cnt += 1;
        }
        cnt = gsum(cnt, red);
        const int kz = (int)cnt;
        if (kz >= N) {
            st |= 1;
        } else if (kz > 0) {
            const double sub = (double)kz * DP / (double)(N - kz);
            for (int j = tid; j < N; j += nt) {
                double v = row[j] - sub;
                row[j] = v < DP ? DP : v;
            }
        }
        double s2 = 0;
        for (int j = tid; j < N; j += nt) s2 += row[j];
        s2 = gsum(s2, red);
        double s3 = 0;
        for (int j = tid; j < N; j += nt) {
            double v = row[j] / s2;
            row[j] = v;
            s3 += v;
        }
        s3 = gsum(s3, red);
        if (!(fabs(s3 - 1.0) <= 1e-8 + 1e-5)) st |= 4;
    }
    if (tid == 0 && st) atomicOr(status, st);
}

template <typename R>
static int estep_gauss_impl(int64_t B, int64_t T, int N, const R *mean, const R *sd, const void *obs, int obs_dtype,
                            const int32_t *lengths, const R *alpha, const R *beta, const R *scale, R *V, double *counts,
                            khmm_stream_t stream) {
    KHMM_REQUIRE(B >= 0 && T >= 1 && N >= 1 && N <= 1024, "estep_gauss: bad shape");
    if (B == 0) return KHMM_OK;
    KHMM_REQUIRE(mean && sd && obs && alpha && beta && scale && V && counts, "estep_gauss: NULL pointer argument");
    KHMM_REQUIRE(obs_dtype == KHMM_F32 || obs_dtype == KHMM_F64, "estep_gauss: obs dtype code %d is not f32/f64", obs_dtype);
    const int threads = ((N + 31) / 32) * 32;
    int64_t cap = (int64_t)sm_count() * (threads <= 128 ? 8 : (threads <= 256 ? 4 : (threads <= 512 ? 2 : 1)));
    int grid = (int)(B < cap ? B : cap);
    gauss_estep_kernel<R><<<grid, threads, 0, (cudaStream_t)stream>>>(B, T, N, mean, sd, obs, obs_dtype, lengths, alpha,
                                                                     beta, scale, V, counts);
    KHMM_LAUNCH_CHECK("gauss_estep_kernel");
    return KHMM_OK;
}

}  // namespace khmm

using namespace khmm;

extern "C" {

size_t khmm_counts_len_gauss(int N) { return (size_t)5 * N + (size_t)N * N + 3; }

int khmm_estep_gauss_f32(int64_t B, int64_t T, int N, const float *mean, const float *sd, const void *obs,
                         int obs_dtype, const int32_t *lengths, const float *alpha, const float *beta,
                         const float *scale, float *V, double *counts, khmm_stream_t stream) {
    return estep_gauss_impl<float>(B, T, N, mean, sd, obs, obs_dtype, lengths, alpha, beta, scale, V, counts, stream);
}
int khmm_estep_gauss_f64(int64_t B, int64_t T, int N, const double *mean, const double *sd, const void *obs,
                         int obs_dtype, const int32_t *lengths, const double *alpha, const double *beta,
                         const double *scale, double *V, double *counts, khmm_stream_t stream) {
    return estep_gauss_impl<double>(B, T, N, mean, sd, obs, obs_dtype, lengths, alpha, beta, scale, V, counts, stream);
}

int khmm_mstep_gauss_f64(int N, const double *counts, const double *A, double sd_min, double *pi_new, double *A_new,
                         double *mean_new, double *sd_new, int32_t *status, khmm_stream_t stream) {
    KHMM_REQUIRE(N >= 1 && counts && A && pi_new && A_new && mean_new && sd_new && status, "mstep_gauss: bad arguments");
    KHMM_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int32_t), (cudaStream_t)stream));
    mstep_gauss_kernel<<<N + 1, 256, 0, (cudaStream_t)stream>>>(N, counts, A, sd_min, pi_new, A_new, mean_new, sd_new,
                                                               status);
    KHMM_LAUNCH_CHECK("mstep_gauss_kernel");
    return KHMM_OK;
}

}  // extern "C"
