// Small elementwise kernels that complete the reference method surface on caller-supplied arrays:
// gamma (abstract_hmm.py:291-297), materialised zi (:299-325), dense emission matrices (:353-367).
#include "common.cuh"

namespace khmm {

template <typename R>
__global__ void gamma_kernel(int64_t n, int N, const R *__restrict__ alpha, const R *__restrict__ beta,
                             const R *__restrict__ scale, R *__restrict__ out, int kind) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        R g = alpha[k] * beta[k];
        if (kind == KHMM_GAMMA_REFERENCE) g = g * scale[k / N];
        out[k] = g;
    }
}

__global__ void zi_kernel(int64_t T, int N, const double *__restrict__ A, const double *__restrict__ E,
                          const double *__restrict__ alpha, const double *__restrict__ beta, double *__restrict__ zi) {
    int64_t n = (T - 1) * (int64_t)N * N;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(k % N);
        int i = (int)((k / N) % N);
        int64_t t = k / ((int64_t)N * N);
        // same left-to-right product order as abstract_hmm.py:316-319
        double v = alpha[t * N + i] * A[i * N + j];
        v = v * E[(t + 1) * N + j];
        v = v * beta[(t + 1) * N + j];
        zi[k] = v;
    }
}

template <class Emis>
__global__ void emissions_kernel(int64_t B, int64_t T, int N, Emis em, double *__restrict__ E) {
    int64_t n = B * T * N;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(k % N);
        int64_t bt = k / N;
        auto o = em.load(bt / T, bt % T);
        E[k] = em.eval(o, j);
    }
}

__global__ void emissions_full_kernel(int64_t rows, int N, int D, const double *__restrict__ means,
                                      const double *__restrict__ Linv, const double *__restrict__ lognorm,
                                      const void *__restrict__ obs, int obs_dtype, double *__restrict__ E) {
    int64_t n = rows * N;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(k % N);
        int64_t r = k / N;
        const double *mu = means + (int64_t)j * D;
        const double *L = Linv + (int64_t)j * D * D;
        double maha = 0.0;
        for (int a = 0; a < D; a++) {
            double y = 0.0;
            for (int b = 0; b <= a; b++) y += L[a * D + b] * (load_real<double>(obs, obs_dtype, r * D + b) - mu[b]);
            maha += y * y;
        }
        E[k] = exp(lognorm[j] - 0.5 * maha);
    }
}

static int ew_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

template <typename R>
static int gamma_impl(int64_t rows, int N, const R *alpha, const R *beta, const R *scale, R *out, int kind,
                      khmm_stream_t stream) {
    KHMM_REQUIRE(rows >= 0 && N >= 1 && alpha && beta && out, "gamma: bad arguments");
    KHMM_REQUIRE(kind == KHMM_GAMMA_POSTERIOR || scale, "gamma: scale is NULL");
    if (rows == 0) return KHMM_OK;
    int64_t n = rows * N;
    gamma_kernel<R><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(n, N, alpha, beta, scale, out, kind);
    KHMM_LAUNCH_CHECK("gamma_kernel");
    return KHMM_OK;
}

}  // namespace khmm

using namespace khmm;

extern "C" {

int khmm_gamma_f32(int64_t rows, int N, const float *alpha, const float *beta, const float *scale, float *out,
                   int gamma_kind, khmm_stream_t stream) {
    return gamma_impl<float>(rows, N, alpha, beta, scale, out, gamma_kind, stream);
}
int khmm_gamma_f64(int64_t rows, int N, const double *alpha, const double *beta, const double *scale, double *out,
                   int gamma_kind, khmm_stream_t stream) {
    return gamma_impl<double>(rows, N, alpha, beta, scale, out, gamma_kind, stream);
}

int khmm_zi_f64(int64_t T, int N, const double *A, const double *E, const double *alpha, const double *beta,
                double *zi, khmm_stream_t stream) {
    KHMM_REQUIRE(T >= 1 && N >= 1 && A && E && alpha && beta && zi, "zi: bad arguments");
    if (T < 2) return KHMM_OK;
    zi_kernel<<<ew_grid((T - 1) * (int64_t)N * N), 256, 0, (cudaStream_t)stream>>>(T, N, A, E, alpha, beta, zi);
    KHMM_LAUNCH_CHECK("zi_kernel");
    return KHMM_OK;
}

int khmm_emissions_discrete_f64(int64_t B, int64_t T, int N, int M, const double *Bsm, const void *obs,
                                int obs_dtype, double *E, khmm_stream_t stream) {
    KHMM_REQUIRE(B >= 0 && T >= 1 && N >= 1 && M >= 1 && Bsm && obs && E, "emissions_discrete: bad arguments");
    KHMM_REQUIRE(obs_dtype >= KHMM_U8 && obs_dtype <= KHMM_I64, "emissions_discrete: obs dtype %d", obs_dtype);
    if (B == 0) return KHMM_OK;
    DiscreteEmis<double> em{Bsm, obs, obs_dtype, N, M, T};
    emissions_kernel<<<ew_grid(B * T * N), 256, 0, (cudaStream_t)stream>>>(B, T, N, em, E);
    KHMM_LAUNCH_CHECK("emissions_kernel");
    return KHMM_OK;
}

int khmm_emissions_gauss_f64(int64_t B, int64_t T, int N, const double *mean, const double *sd, const void *obs,
                             int obs_dtype, double *E, khmm_stream_t stream) {
    KHMM_REQUIRE(B >= 0 && T >= 1 && N >= 1 && mean && sd && obs && E, "emissions_gauss: bad arguments");
    KHMM_REQUIRE(obs_dtype == KHMM_F32 || obs_dtype == KHMM_F64, "emissions_gauss: obs dtype %d", obs_dtype);
    if (B == 0) return KHMM_OK;
    GaussEmis<double> em{mean, sd, obs, obs_dtype, N, T};
    emissions_kernel<<<ew_grid(B * T * N), 256, 0, (cudaStream_t)stream>>>(B, T, N, em, E);
    KHMM_LAUNCH_CHECK("emissions_kernel");
    return KHMM_OK;
}

int khmm_emissions_gauss_full_f64(int64_t B, int64_t T, int N, int D, const double *means, const double *Linv,
                                  const double *lognorm, const void *obs, int obs_dtype, double *E,
                                  khmm_stream_t stream) {
    KHMM_REQUIRE(B >= 0 && T >= 1 && N >= 1 && D >= 1 && means && Linv && lognorm && obs && E,
                 "emissions_gauss_full: bad arguments");
    KHMM_REQUIRE(obs_dtype == KHMM_F32 || obs_dtype == KHMM_F64, "emissions_gauss_full: obs dtype %d", obs_dtype);
    if (B == 0) return KHMM_OK;
    emissions_full_kernel<<<ew_grid(B * T * N), 256, 0, (cudaStream_t)stream>>>(B * T, N, D, means, Linv, lognorm, obs,
                                                                               obs_dtype, E);
    KHMM_LAUNCH_CHECK("emissions_full_kernel");
    return KHMM_OK;
}

}  // extern "C"
