// Batched ancestral sampling -- AbstractHMM.simulate / transition / emit (abstract_hmm.py:444-469,
// distribution.py:38-39, :157-161) for a whole batch of independent sequences.
//
// Per step, exactly like the reference: transition first (from pi when there is no current state, else
// from row A[current]), then emit from the new state's distribution.  numpy.random.choice(p=...) is
// "cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, side='right')" with u ~ U[0,1): the caller
// passes those float64 CDFs (built on the host with the same NumPy calls) and the kernel does the same
// right-sided binary search, so that with identical uniforms the states and symbols are identical.
// Gaussian emissions are mean + sd * z (scipy.stats.norm(loc=mean, scale=variance).rvs(), the
// reference's `variance` is the scale) with z from the Box-Muller transform in float64.
//
// Uniforms come from Philox4x32-10 (counter-based: counter = (sequence index, step), key = seed), so a
// sequence's stream does not depend on the batch shape or launch geometry -- or from a caller-supplied
// array (tests pin the sampling logic that way; the reference's own RNG stream, NumPy's global
// MT19937, cannot be reproduced on the device: parity of the values is distributional only).
// One thread per sequence; this path generates synthetic training data (5.4e8 symbols at config 4), it
// is bandwidth-trivial and not a bench line.
#include "common.cuh"

namespace khmm {

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ static void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
        uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    __device__ __forceinline__ void operator()(uint32_t (&c)[4]) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            round(c, a, b);
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
    }
};

// 53-bit uniform in [0, 1) from two 32-bit words (the construction NumPy's random() uses)
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return (double)(((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

// numpy.searchsorted(cdf[0:n], u, side="right"), clamped to n-1 (cdf[n-1] == 1 > u always holds for a
// normalised CDF; the clamp only guards a CDF whose last entry rounds below 1)
__device__ __forceinline__ int search_right(const double *__restrict__ cdf, int n, double u) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(cdf + mid) <= u) lo = mid + 1; else hi = mid;
    }
    return lo < n ? lo : n - 1;
}

template <bool GAUSS>
__global__ void __launch_bounds__(128)
simulate_kernel(int64_t B, int64_t T, int N, int M, const double *__restrict__ cum_pi, const double *__restrict__ cumA,
                const double *__restrict__ cumB, const double *__restrict__ mean, const double *__restrict__ sd,
                uint64_t seed, const double *__restrict__ uniforms, int32_t *__restrict__ states,
                int32_t *__restrict__ symbols, double *__restrict__ values) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const Philox rng{(uint32_t)seed, (uint32_t)(seed >> 32)};
    int cur = -1;
    for (int64_t t = 0; t < T; t++) {
        double u_state, u_a, u_b;
        if (uniforms) {
            const double *u = uniforms + (b * T + t) * 3;
            u_state = u[0]; u_a = u[1]; u_b = u[2];
        } else {
            uint32_t c[4] = {(uint32_t)b, (uint32_t)(b >> 32), (uint32_t)t, (uint32_t)(t >> 32)};
            rng(c);
            u_state = u53(c[0], c[1]);
            u_a = u53(c[2], c[3]);
            u_b = 0.0;
            if (GAUSS) {   // a second block for the second Box-Muller uniform (counter word 3 gets bit 31)
                uint32_t d[4] = {(uint32_t)b, (uint32_t)(b >> 32), (uint32_t)t, (uint32_t)(t >> 32) | 0x80000000u};
                rng(d);
                u_b = u53(d[0], d[1]);
            }
        }
        cur = search_right(cur < 0 ? cum_pi : cumA + (int64_t)cur * N, N, u_state);   // transition(), :449-456
        states[b * T + t] = cur;
        if (GAUSS) {                                                                   // emit(), :444-447
            // Box-Muller; u_a in [0,1) -> 1-u_a in (0,1] keeps the logarithm finite
            double z = sqrt(-2.0 * log(1.0 - u_a)) * cospi(2.0 * u_b);
            values[b * T + t] = __ldg(mean + cur) + __ldg(sd + cur) * z;
        } else {
            symbols[b * T + t] = search_right(cumB + (int64_t)cur * M, M, u_a);
        }
    }
}

__global__ void philox_kat_kernel(const uint32_t *__restrict__ ctr_key, uint32_t *__restrict__ out, int n) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t c[4] = {ctr_key[6 * k], ctr_key[6 * k + 1], ctr_key[6 * k + 2], ctr_key[6 * k + 3]};
    Philox rng{ctr_key[6 * k + 4], ctr_key[6 * k + 5]};
    rng(c);
    for (int q = 0; q < 4; q++) out[4 * k + q] = c[q];
}

}  // namespace khmm

using namespace khmm;

extern "C" {

int khmm_simulate_discrete(int64_t B, int64_t T, int N, int M, const double *cum_pi, const double *cumA,
                           const double *cumB, uint64_t seed, const double *uniforms, int32_t *states,
                           int32_t *symbols, khmm_stream_t stream) {
    KHMM_REQUIRE(B >= 0 && T >= 1 && N >= 1 && M >= 1, "simulate_discrete: bad shape");
    if (B == 0) return KHMM_OK;
    KHMM_REQUIRE(cum_pi && cumA && cumB && states && symbols, "simulate_discrete: NULL pointer argument");
    simulate_kernel<false><<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        B, T, N, M, cum_pi, cumA, cumB, nullptr, nullptr, seed, uniforms, states, symbols, nullptr);
    KHMM_LAUNCH_CHECK("simulate_kernel");
    return KHMM_OK;
}

int khmm_simulate_gauss(int64_t B, int64_t T, int N, const double *cum_pi, const double *cumA, const double *mean,
                        const double *sd, uint64_t seed, const double *uniforms, int32_t *states, double *values,
                        khmm_stream_t stream) {
    KHMM_REQUIRE(B >= 0 && T >= 1 && N >= 1, "simulate_gauss: bad shape");
    if (B == 0) return KHMM_OK;
    KHMM_REQUIRE(cum_pi && cumA && mean && sd && states && values, "simulate_gauss: NULL pointer argument");
    simulate_kernel<true><<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        B, T, N, 0, cum_pi, cumA, nullptr, mean, sd, seed, uniforms, states, nullptr, values);
    KHMM_LAUNCH_CHECK("simulate_kernel");
    return KHMM_OK;
}

int khmm_philox4x32_10(int n, const uint32_t *ctr_key, uint32_t *out, khmm_stream_t stream) {
    KHMM_REQUIRE(n >= 1 && ctr_key && out, "philox4x32_10: bad arguments");
    philox_kat_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(ctr_key, out, n);
    KHMM_LAUNCH_CHECK("philox_kat_kernel");
    return KHMM_OK;
}

}  // extern "C"
