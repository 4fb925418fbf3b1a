// Shared device/host helpers for libkhmm (sm_100a).  See include/khmm.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/khmm.h"

namespace khmm {

// ------------------------------------------------------------------ error plumbing
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define KHMM_CUDA_CHECK(expr)                                    \
    do {                                                         \
        cudaError_t _e = (expr);                                 \
        if (_e != cudaSuccess) return khmm::cuda_fail(_e, #expr); \
    } while (0)

#define KHMM_LAUNCH_CHECK(name)                                       \
    do {                                                              \
        cudaError_t _e = cudaGetLastError();                          \
        if (_e != cudaSuccess) return khmm::cuda_fail(_e, name);      \
    } while (0)

#define KHMM_REQUIRE(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            khmm::set_error(__VA_ARGS__);    \
            return KHMM_ERR_INVALID;         \
        }                                    \
    } while (0)

int sm_count();  // cached SM count of the current device (148 on B200)

// Development switches (khmm_debug_switch; the environment variable of the same name gives the initial value, read
// ONCE per process): which kernel variant a call takes.  None is needed in production.
enum DevSwitch { SW_RC_MODE = 0, SW_RC_BN = 1, SW_VIT_SCREEN = 2, SW_BW_BACKWARD = 3, SW_SMALL_MMA = 4, SW_COUNT = 5 };
int dev_switch(DevSwitch which);   // -1 = unset (the library's default applies)

// ---------------------------------------------------------------- observation loads
__device__ __forceinline__ int load_symbol(const void *p, int dtype, int64_t idx) {
    switch (dtype) {
        case KHMM_U8: return (int)((const uint8_t *)p)[idx];
        case KHMM_U16: return (int)((const uint16_t *)p)[idx];
        case KHMM_I32: return ((const int32_t *)p)[idx];
        default: return (int)((const int64_t *)p)[idx];
    }
}
template <typename R>
__device__ __forceinline__ R load_real(const void *p, int dtype, int64_t idx) {
    return dtype == KHMM_F32 ? (R)((const float *)p)[idx] : (R)((const double *)p)[idx];
}

// ---------------------------------------------------------------- emission functors
// eval(b, t, j) = b_j(o_{b,t}); `prep(j)` caches per-state constants when a thread owns state j.
template <typename R>
struct DiscreteEmis {
    const R *Bsm;  // [M][N]
    const void *obs;
    int obs_dtype;
    int N, M;
    int64_t T;
    struct Obs { int sym; };
    __device__ __forceinline__ Obs load(int64_t b, int64_t t) const {
        int s = load_symbol(obs, obs_dtype, b * T + t);
        // out-of-alphabet symbols index nothing valid in the reference either (IndexError);
        // clamp so a bad pad value cannot fault the kernel.
        s = s < 0 ? 0 : (s >= M ? M - 1 : s);
        return Obs{s};
    }
    __device__ __forceinline__ R eval(const Obs &o, int j) const { return __ldg(Bsm + (int64_t)o.sym * N + j); }
};

template <typename R>
struct GaussEmis {
    const R *mean, *sd;
    const void *obs;
    int obs_dtype;
    int N;
    int64_t T;
    struct Obs { R x; };
    __device__ __forceinline__ Obs load(int64_t b, int64_t t) const {
        return Obs{load_real<R>(obs, obs_dtype, b * T + t)};
    }
    // scipy.stats.norm(loc, scale).pdf(x) = exp(-z^2/2)/sqrt(2*pi)/scale, z=(x-loc)/scale
    __device__ __forceinline__ R eval(const Obs &o, int j) const {
        R s = __ldg(sd + j);
        R z = (o.x - __ldg(mean + j)) / s;
        if (sizeof(R) == 8)
            return (R)(exp(-0.5 * (double)z * (double)z) * 0.3989422804014327 / (double)s);
        return (R)(expf(-0.5f * (float)z * (float)z) * 0.3989422804014327f / (float)s);
    }
};

// Full-covariance Gaussian, D <= KHMM_GAUSS_FULL_DMAX features (distribution.py:153: scipy multivariate_normal(mean, cov).pdf):
// b_j(x) = exp(lognorm_j - 0.5 |Linv_j (x - mean_j)|^2), Linv_j = inverse of the lower Cholesky factor of cov_j, lognorm_j =
// -0.5 (D log 2 pi + log det cov_j).  Evaluated in float64 whatever R, operation for operation like
// khmm_emissions_gauss_full_f64 (so the fused path and the dense-matrix path give the same emission values), then cast.
template <typename R>
struct GaussFullEmis {
    const double *means, *Linv, *lognorm;   // [N][D], [N][D][D] (row-major, lower triangular), [N]
    const void *obs;                        // [B][T][D]
    int obs_dtype;
    int N, D;
    int64_t T;
    struct Obs { double x[KHMM_GAUSS_FULL_DMAX]; };
    __device__ __forceinline__ Obs load(int64_t b, int64_t t) const {
        Obs o;
#pragma unroll
        for (int a = 0; a < KHMM_GAUSS_FULL_DMAX; a++) o.x[a] = a < D ? load_real<double>(obs, obs_dtype, (b * T + t) * D + a) : 0.0;
        return o;
    }
    __device__ __forceinline__ R eval(const Obs &o, int j) const {
        const double *mu = means + (int64_t)j * D;
        const double *L = Linv + (int64_t)j * D * D;
        double maha = 0.0;
#pragma unroll
        for (int a = 0; a < KHMM_GAUSS_FULL_DMAX; a++) {
            if (a < D) {
                double y = 0.0;
#pragma unroll
                for (int c = 0; c < KHMM_GAUSS_FULL_DMAX; c++)
                    if (c <= a) y += __ldg(L + a * D + c) * (o.x[c] - __ldg(mu + c));
                maha += y * y;
            }
        }
        return (R)exp(__ldg(lognorm + j) - 0.5 * maha);
    }
};

template <typename R>
struct DenseEmis {
    const R *E;  // [B][T][N]
    int N;
    int64_t T;
    struct Obs { const R *row; };
    __device__ __forceinline__ Obs load(int64_t b, int64_t t) const { return Obs{E + (b * T + t) * N}; }
    __device__ __forceinline__ R eval(const Obs &o, int j) const { return __ldg(o.row + j); }
};

// ------------------------------------------------------------------- warp reductions
template <typename R>
__device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace khmm
