// Device-resident Baum-Welch loop helpers: the re-estimated parameters of one iteration become the operands of the
// next without leaving the GPU (AbstractHMM.train, abstract_hmm.py:400-441, calls DiscreteHMM.do_pass once per
// iteration and do_pass installs pi / A / B at discrete_hmm.py:168-173 -- on the host, in the reference).
//
//   commit_discrete_kernel   installs the M-step results (float64, state-major B) as the master parameters -- unless
//       this or an earlier iteration hit one of the reference's exceptional cases (status bits of
//       khmm_mstep_discrete_f64), in which case the parameters of the last good iteration stay and the first failing
//       iteration is remembered -- and writes the operand copies the kernels read: pi, A, A^T and the SYMBOL-MAJOR
//       emission table, in float32 or float64, optionally padded to Np >= N states with unreachable states
//       (start probability 0, no transitions, zero emission row).
//   symbol_range_kernel      min / max of the live symbols of a batch (the reference raises IndexError for a symbol
//       >= M and wraps negative ones, distribution.py:35-36; the kernels only clamp, so the host checks once).
#include "common.cuh"

namespace khmm {

template <typename R>
__global__ void __launch_bounds__(256)
commit_discrete_kernel(int N, int M, int Np, int iteration, const int32_t *__restrict__ status_it,
                       int32_t *__restrict__ sticky, const double *__restrict__ pi_src, const double *__restrict__ A_src,
                       const double *__restrict__ B_src, double *__restrict__ pi64, double *__restrict__ A64,
                       double *__restrict__ B64, R *__restrict__ piR, R *__restrict__ AR, R *__restrict__ ATR,
                       R *__restrict__ BsmR) {
    const int st = status_it ? *status_it : 0;
    const int old = sticky ? sticky[0] : 0;
    const bool fail = (st != 0) || (old != 0);
    // every block reads the flags before block 0 may update them: the decision `fail` is the same either way
    // (st != 0 whenever the update happens)
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0 && sticky && st != 0 && old == 0) {
        sticky[0] = st;
        sticky[1] = iteration + 1;
    }
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    // 1. masters (skipped when the sources ARE the masters: the initial pack)
    const bool copy = !fail && pi_src != pi64;
    // 2. operand copies always come from what the masters hold after this kernel
    const double *pi_v = fail ? pi64 : pi_src;
    const double *A_v = fail ? A64 : A_src;
    const double *B_v = fail ? B64 : B_src;
    for (int64_t k = tid; k < Np; k += nth) {
        double v = k < N ? pi_v[k] : 0.0;
        if (copy && k < N) pi64[k] = v;
        piR[k] = (R)v;
    }
    for (int64_t k = tid; k < (int64_t)Np * Np; k += nth) {
        const int i = (int)(k / Np), j = (int)(k % Np);
        double v = (i < N && j < N) ? A_v[(int64_t)i * N + j] : 0.0;
        if (copy && i < N && j < N) A64[(int64_t)i * N + j] = v;
        AR[k] = (R)v;
        ATR[(int64_t)j * Np + i] = (R)v;
    }
    // emission table: read state-major [N][M] along k (coalesced), write symbol-major [M][Np]
    for (int64_t k = tid; k < (int64_t)Np * M; k += nth) {
        const int i = (int)(k / M), s = (int)(k % M);
        double v = i < N ? B_v[(int64_t)i * M + s] : 0.0;
        if (copy && i < N) B64[(int64_t)i * M + s] = v;
        BsmR[(int64_t)s * Np + i] = (R)v;
    }
}

__global__ void __launch_bounds__(256)
symbol_range_kernel(int64_t B, int64_t T, const void *__restrict__ obs, int dtype, const int32_t *__restrict__ lengths,
                    long long *__restrict__ out) {
    long long lo = 0x7fffffffffffffffLL, hi = -0x7fffffffffffffffLL - 1;
    const int64_t n = B * T;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        if (lengths) {
            const int64_t b = k / T;
            if (k - b * T >= lengths[b]) continue;
        }
        long long s;
        switch (dtype) {
            case KHMM_U8: s = ((const uint8_t *)obs)[k]; break;
            case KHMM_U16: s = ((const uint16_t *)obs)[k]; break;
            case KHMM_I32: s = ((const int32_t *)obs)[k]; break;
            default: s = ((const int64_t *)obs)[k]; break;
        }
        lo = s < lo ? s : lo;
        hi = s > hi ? s : hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, lo);
        atomicMax(out + 1, hi);
    }
}

template <typename R>
static int commit_impl(int N, int M, int Np, int iteration, const int32_t *status_it, int32_t *sticky,
                       const double *pi_src, const double *A_src, const double *B_src, double *pi64, double *A64,
                       double *B64, R *piR, R *AR, R *ATR, R *BsmR, khmm_stream_t stream) {
    KHMM_REQUIRE(N >= 1 && M >= 1 && Np >= N, "train_commit: bad shape");
    KHMM_REQUIRE(pi_src && A_src && B_src && pi64 && A64 && B64 && piR && AR && ATR && BsmR,
                 "train_commit: NULL pointer argument");
    KHMM_REQUIRE((pi_src == pi64) == (A_src == A64) && (A_src == A64) == (B_src == B64),
                 "train_commit: either all sources alias the masters (initial pack) or none does");
    int64_t work = (int64_t)Np * (M > Np ? M : Np);
    int grid = (int)((work + 255) / 256);
    grid = grid < 1 ? 1 : (grid > 4 * sm_count() ? 4 * sm_count() : grid);
    commit_discrete_kernel<R><<<grid, 256, 0, (cudaStream_t)stream>>>(N, M, Np, iteration, status_it, sticky, pi_src,
                                                                     A_src, B_src, pi64, A64, B64, piR, AR, ATR, BsmR);
    KHMM_LAUNCH_CHECK("commit_discrete_kernel");
    return KHMM_OK;
}

}  // namespace khmm

using namespace khmm;

extern "C" {

int khmm_train_commit_discrete_f32(int N, int M, int Np, int iteration, const int32_t *status_it, int32_t *sticky,
                                   const double *pi_src, const double *A_src, const double *B_src, double *pi64,
                                   double *A64, double *B64, float *pi_op, float *A_op, float *AT_op, float *Bsm_op,
                                   khmm_stream_t stream) {
    return commit_impl<float>(N, M, Np, iteration, status_it, sticky, pi_src, A_src, B_src, pi64, A64, B64, pi_op, A_op,
                              AT_op, Bsm_op, stream);
}
int khmm_train_commit_discrete_f64(int N, int M, int Np, int iteration, const int32_t *status_it, int32_t *sticky,
                                   const double *pi_src, const double *A_src, const double *B_src, double *pi64,
                                   double *A64, double *B64, double *pi_op, double *A_op, double *AT_op, double *Bsm_op,
                                   khmm_stream_t stream) {
    return commit_impl<double>(N, M, Np, iteration, status_it, sticky, pi_src, A_src, B_src, pi64, A64, B64, pi_op, A_op,
                               AT_op, Bsm_op, stream);
}

int khmm_symbol_range(int64_t B, int64_t T, const void *obs, int obs_dtype, const int32_t *lengths, int64_t *min_max,
                      khmm_stream_t stream) {
    KHMM_REQUIRE(B >= 0 && T >= 1 && obs && min_max, "symbol_range: bad arguments");
    KHMM_REQUIRE(obs_dtype >= KHMM_U8 && obs_dtype <= KHMM_I64, "symbol_range: obs dtype code %d is not an integer type",
                 obs_dtype);
    const long long init[2] = {0x7fffffffffffffffLL, -0x7fffffffffffffffLL - 1};
    KHMM_CUDA_CHECK(cudaMemcpyAsync(min_max, init, sizeof(init), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    if (B == 0) return KHMM_OK;
    int64_t n = B * T;
    int grid = (int)((n + 256 * 16 - 1) / (256 * 16));
    grid = grid < 1 ? 1 : (grid > 8 * sm_count() ? 8 * sm_count() : grid);
    symbol_range_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(B, T, obs, obs_dtype, lengths, (long long *)min_max);
    KHMM_LAUNCH_CHECK("symbol_range_kernel");
    return KHMM_OK;
}

}  // extern "C"
