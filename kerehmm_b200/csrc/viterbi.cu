// Viterbi: float64 max-plus DP with first-index argmax (bit-exact against
// AbstractHMM.viterbi_path, kerehmm/abstract_hmm.py:203-239) + backtrace.
//
// Bit-exactness contract: the log tables are computed on the HOST with
// np.log and only added/compared here, in the reference's association order:
//     transitions_i = delta[t-1][i] + logA[i][j]      (one IEEE add,   abstract_hmm.py:227)
//     m = max_i, psi = first argmax_i                  (compares only,  :228,:230)
//     delta[t][j] = m + logB_j(o_t)                    (one IEEE add,   :228-229)
// No FMA contraction is possible (there is no multiply) and no log is taken on the device.
//
// Generic kernel: thread j owns target state j of S sequences; delta_{t-1} sits in shared memory
// as [N][S] (broadcast vector loads), logA is read [i][j] (coalesced / conflict free) from shared
// memory when it fits, else through L1/L2.  Backpointers go to HBM as uint8 (N <= 256) or uint16,
// one coalesced N-byte row per (sequence, step).
#include "common.cuh"

namespace khmm {

constexpr int kVitSeq = 4;

template <typename PSI, bool A_SMEM>
__global__ void __launch_bounds__(1024)
viterbi_generic_kernel(int64_t B, int64_t T, int N, int M, const double *__restrict__ logpi,
                       const double *__restrict__ logA, const double *__restrict__ logBsm, const void *__restrict__ obs,
                       int obs_dtype, const int32_t *__restrict__ lengths, PSI *__restrict__ psi,
                       int32_t *__restrict__ last_state, double *__restrict__ logp) {
    constexpr int S = kVitSeq;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = blockDim.x, j = threadIdx.x;
    double *rows = reinterpret_cast<double *>(smem_raw);  // [2][NT][S]
    double *As = rows + 2 * NT * S;
    const double *Am = logA;
    if (A_SMEM) {
        for (int k = threadIdx.x; k < N * N; k += blockDim.x) As[k] = logA[k];
        Am = As;
    }
    const bool active = j < N;
    const double lpj = active ? logpi[j] : 0.0;
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
        int cur = 0;
        for (int t = 0; t < tmax; t++) {
            const double *prev = rows + cur * NT * S;
            double *next = rows + (cur ^ 1) * NT * S;
            double best[S];
            int arg[S];
            if (t == 0) {
#pragma unroll
                for (int s = 0; s < S; s++) { best[s] = lpj; arg[s] = 0; }
            } else if (active) {
                {
                    double a0 = Am[j];
                    double2 p0 = *reinterpret_cast<const double2 *>(prev);
                    double2 p1 = *reinterpret_cast<const double2 *>(prev + 2);
                    best[0] = p0.x + a0; best[1] = p0.y + a0; best[2] = p1.x + a0; best[3] = p1.y + a0;
#pragma unroll
                    for (int s = 0; s < S; s++) arg[s] = 0;
                }
#pragma unroll 4
                for (int i = 1; i < N; i++) {
                    double aij = Am[i * N + j];
                    double2 p0 = *reinterpret_cast<const double2 *>(prev + i * S);
                    double2 p1 = *reinterpret_cast<const double2 *>(prev + i * S + 2);
                    double v[S] = {p0.x + aij, p0.y + aij, p1.x + aij, p1.y + aij};
#pragma unroll
                    for (int s = 0; s < S; s++)
                        if (v[s] > best[s]) { best[s] = v[s]; arg[s] = i; }
                }
            }
#pragma unroll
            for (int s = 0; s < S; s++) {
                double d = 0.0;
                if (active && t < len[s]) {
                    int64_t b = b0 + s;
                    int sym = load_symbol(obs, obs_dtype, b * T + t);
                    sym = sym < 0 ? 0 : (sym >= M ? M - 1 : sym);
                    d = best[s] + __ldg(logBsm + (int64_t)sym * N + j);
                    if (t > 0) psi[(b * T + t) * N + j] = (PSI)arg[s];
                }
                next[j * S + s] = d;
            }
            __syncthreads();
            // termination for sequences that end at this step: first-index argmax (abstract_hmm.py:233,235)
#pragma unroll
            for (int s = 0; s < S; s++) {
                if (j == s && t == len[s] - 1) {
                    double m = next[s];
                    int am = 0;
                    for (int k = 1; k < N; k++) {
                        double v = next[k * S + s];
                        if (v > m) { m = v; am = k; }
                    }
                    last_state[b0 + s] = am;
                    logp[b0 + s] = m;
                }
            }
            cur ^= 1;
        }
    }
}

// ------------------------------------------------------------------ register-tile kernel
// N in {32, 64, 128}: logA lives in REGISTERS.  A CTA has W = N*N/2048 warps; warp w owns the
// sources [w*SRC, (w+1)*SRC) with SRC = N/W, lane l owns the JT = N/32 targets l, l+32, ...: a
// JT x SRC tile of logA (64 doubles = 128 registers).  Each delta value is fetched from shared memory
// once per JT transitions as a pure broadcast, so the inner step is DADD + DSETP + 3 selects with no
// memory traffic of its own (measured 1.5x the generic kernel at N=64; tools/vit64_bench.cu holds
// the experiment).  Warps exchange their partial (best, arg) through shared memory once per
// (sequence, step); the first N threads merge them in source order -- a later source range wins
// only when strictly greater, which keeps np.argmax's first-index rule.
template <int NN, int S>
struct VitTile {
    static constexpr int W = NN * NN / 2048 < 1 ? 1 : NN * NN / 2048;
    static constexpr int SRC = NN / W;
    static constexpr int JT = NN / 32;
    static constexpr int THREADS = 32 * W;
    static constexpr size_t smem_bytes() {
        return sizeof(double) * (2 * S * NN + S * W * NN) + sizeof(int) * S * W * NN;
    }
};

template <int NN, int S, typename PSI>
__global__ void __launch_bounds__(VitTile<NN, S>::THREADS)
viterbi_tile_kernel(int64_t B, int64_t T, int M, const double *__restrict__ logpi, const double *__restrict__ logA,
                    const double *__restrict__ logBsm, const void *__restrict__ obs, int obs_dtype,
                    const int32_t *__restrict__ lengths, PSI *__restrict__ psi, int32_t *__restrict__ last_state,
                    double *__restrict__ logp) {
    using C = VitTile<NN, S>;
    constexpr int W = C::W, SRC = C::SRC, JT = C::JT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *d = reinterpret_cast<double *>(smem_raw);          // [2][S][NN]
    double *pbest = d + 2 * S * NN;                            // [S][W][NN]
    int *parg = reinterpret_cast<int *>(pbest + S * W * NN);   // [S][W][NN]
    const int tid = threadIdx.x, l = tid & 31, w = tid >> 5;
    double a[JT][SRC];
#pragma unroll
    for (int q = 0; q < JT; q++)
#pragma unroll
        for (int k = 0; k < SRC; k++) a[q][k] = logA[(SRC * w + k) * NN + l + 32 * q];
    const bool owner = tid < NN;  // threads that own a target state for merging / emission / output
    const double lp = owner ? logpi[tid] : 0.0;
    const int64_t ngroups = (B + S - 1) / S;
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
        int cur = 0;
        for (int t = 0; t < tmax; t++) {
            double e[S];
            if (owner) {
#pragma unroll
                for (int s = 0; s < S; s++) {
                    e[s] = 0.0;
                    if (t < len[s]) {
                        int sym = load_symbol(obs, obs_dtype, (b0 + s) * T + t);
                        sym = sym < 0 ? 0 : (sym >= M ? M - 1 : sym);
                        e[s] = __ldg(logBsm + (int64_t)sym * NN + tid);
                    }
                }
            }
            if (t > 0) {
#pragma unroll
                for (int s = 0; s < S; s++) {
                    const double *dp = d + (cur * S + s) * NN + SRC * w;
                    double best[JT];
                    int arg[JT];
                    {
                        const double dv = dp[0];
#pragma unroll
                        for (int q = 0; q < JT; q++) { best[q] = dv + a[q][0]; arg[q] = 0; }
                    }
#pragma unroll
                    for (int k = 1; k < SRC; k++) {
                        const double dv = dp[k];
#pragma unroll
                        for (int q = 0; q < JT; q++) {
                            const double v = dv + a[q][k];
                            if (v > best[q]) { best[q] = v; arg[q] = k; }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < JT; q++) {
                        pbest[(s * W + w) * NN + l + 32 * q] = best[q];
                        parg[(s * W + w) * NN + l + 32 * q] = arg[q] + SRC * w;
                    }
                }
                __syncthreads();
            }
            if (owner) {
#pragma unroll
                for (int s = 0; s < S; s++) {
                    double nd = 0.0;
                    if (t < len[s]) {
                        if (t == 0) {
                            nd = lp + e[s];
                        } else {
                            double bb = pbest[(s * W) * NN + tid];
                            int ba = parg[(s * W) * NN + tid];
#pragma unroll
                            for (int q = 1; q < W; q++) {
                                double qb = pbest[(s * W + q) * NN + tid];
                                int qa = parg[(s * W + q) * NN + tid];
                                if (qb > bb) { bb = qb; ba = qa; }
                            }
                            psi[((b0 + s) * T + t) * NN + tid] = (PSI)ba;
                            nd = bb + e[s];
                        }
                    }
                    d[((cur ^ 1) * S + s) * NN + tid] = nd;
                }
            }
            __syncthreads();
            cur ^= 1;
            // termination: first-index argmax over the final delta row (abstract_hmm.py:233,235)
#pragma unroll
            for (int s = 0; s < S; s++) {
                if (tid == s && t == len[s] - 1) {
                    const double *row = d + (cur * S + s) * NN;
                    double m = row[0];
                    int am = 0;
                    for (int k = 1; k < NN; k++)
                        if (row[k] > m) { m = row[k]; am = k; }
                    last_state[b0 + s] = am;
                    logp[b0 + s] = m;
                }
            }
        }
        __syncthreads();
    }
}

template <int NN, int S, typename PSI>
static int launch_viterbi_tile(int64_t B, int64_t T, int M, const double *logpi, const double *logA,
                               const double *logBsm, const void *obs, int obs_dtype, const int32_t *lengths, PSI *psi,
                               int32_t *last, double *logp, cudaStream_t st) {
    using C = VitTile<NN, S>;
    auto kern = viterbi_tile_kernel<NN, S, PSI>;
    size_t smem = C::smem_bytes();
    if (smem > 48 * 1024)
        KHMM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    KHMM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::THREADS, smem));
    if (occ < 1) occ = 1;
    int64_t ngroups = (B + S - 1) / S;
    int64_t cap = (int64_t)sm_count() * occ;   // persistent: one resident wave, grid-stride over groups
    int grid = (int)(ngroups < cap ? ngroups : cap);
    kern<<<grid, C::THREADS, smem, st>>>(B, T, M, logpi, logA, logBsm, obs, obs_dtype, lengths, psi, last, logp);
    KHMM_LAUNCH_CHECK("viterbi_tile_kernel");
    return KHMM_OK;
}

// One thread per sequence chases the backpointers: path[t] = psi[t+1][path[t+1]] (:236-237).
template <typename PSI, typename OUT>
__global__ void viterbi_backtrace_kernel(int64_t B, int64_t T, int N, const int32_t *__restrict__ lengths,
                                         const PSI *__restrict__ psi, const int32_t *__restrict__ last_state,
                                         OUT *__restrict__ path) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int len = lengths ? lengths[b] : (int)T;
    int st = last_state[b];
    OUT *row = path + b * T;
    const PSI *prow = psi + b * T * N;
    row[len - 1] = (OUT)st;
    for (int t = len - 2; t >= 0; t--) {
        st = (int)prow[(int64_t)(t + 1) * N + st];
        row[t] = (OUT)st;
    }
}

template <typename PSI>
static int run_viterbi(int64_t B, int64_t T, int N, int M, const double *logpi, const double *logA,
                       const double *logBsm, const void *obs, int obs_dtype, const int32_t *lengths, void *backptr,
                       void *path, int path_dtype, double *logp, cudaStream_t st) {
    constexpr int S = kVitSeq;
    int NT = ((N + 31) / 32) * 32;
    size_t rows_b = sizeof(double) * 2 * NT * S;
    size_t a_b = sizeof(double) * (size_t)N * N;
    bool as = rows_b + a_b <= 200 * 1024;
    size_t smem = rows_b + (as ? a_b : 0);
    // the first B int32 of the workspace hold the terminal states, psi follows (256-byte aligned)
    int32_t *last = reinterpret_cast<int32_t *>(backptr);
    size_t off = (((size_t)B * sizeof(int32_t)) + 255) / 256 * 256;
    PSI *psi = reinterpret_cast<PSI *>(reinterpret_cast<unsigned char *>(backptr) + off);
    int rc_tile = -1;
    if (sizeof(PSI) == 1 && B >= 64) {   // register-tile kernels (small batches stay on the generic kernel)
        if (N == 32) rc_tile = launch_viterbi_tile<32, 4, PSI>(B, T, M, logpi, logA, logBsm, obs, obs_dtype, lengths, psi, last, logp, st);
        else if (N == 64) rc_tile = launch_viterbi_tile<64, 4, PSI>(B, T, M, logpi, logA, logBsm, obs, obs_dtype, lengths, psi, last, logp, st);
        else if (N == 128) rc_tile = launch_viterbi_tile<128, 2, PSI>(B, T, M, logpi, logA, logBsm, obs, obs_dtype, lengths, psi, last, logp, st);
    }
    if (rc_tile > 0) return rc_tile;
    if (rc_tile < 0) {
        int per_sm = NT <= 64 ? 16 : (NT <= 128 ? 8 : (NT <= 256 ? 4 : (NT <= 512 ? 2 : 1)));
        if (smem > 48 * 1024) per_sm = (int)max((size_t)1, min((size_t)per_sm, (size_t)(220 * 1024) / smem));
        int64_t ngroups = (B + S - 1) / S;
        int64_t cap = (int64_t)sm_count() * per_sm;
        int grid = (int)(ngroups < cap ? ngroups : cap);
        auto kern = as ? viterbi_generic_kernel<PSI, true> : viterbi_generic_kernel<PSI, false>;
        if (smem > 48 * 1024)
            KHMM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, NT, smem, st>>>(B, T, N, M, logpi, logA, logBsm, obs, obs_dtype, lengths, psi, last, logp);
        KHMM_LAUNCH_CHECK("viterbi_generic_kernel");
    }
    int bt = 128;
    int bg = (int)((B + bt - 1) / bt);
    switch (path_dtype) {
        case KHMM_U8: viterbi_backtrace_kernel<PSI, uint8_t><<<bg, bt, 0, st>>>(B, T, N, lengths, psi, last, (uint8_t *)path); break;
        case KHMM_U16: viterbi_backtrace_kernel<PSI, uint16_t><<<bg, bt, 0, st>>>(B, T, N, lengths, psi, last, (uint16_t *)path); break;
        case KHMM_I32: viterbi_backtrace_kernel<PSI, int32_t><<<bg, bt, 0, st>>>(B, T, N, lengths, psi, last, (int32_t *)path); break;
        default: viterbi_backtrace_kernel<PSI, int64_t><<<bg, bt, 0, st>>>(B, T, N, lengths, psi, last, (int64_t *)path); break;
    }
    KHMM_LAUNCH_CHECK("viterbi_backtrace_kernel");
    return KHMM_OK;
}

}  // namespace khmm

using namespace khmm;

extern "C" {

size_t khmm_viterbi_workspace_bytes(int64_t B, int64_t T, int N) {
    size_t off = (((size_t)B * sizeof(int32_t)) + 255) / 256 * 256;
    size_t w = N <= 256 ? 1 : 2;
    return off + (size_t)B * (size_t)T * (size_t)N * w + 256;
}

int khmm_viterbi_discrete_f64(int64_t B, int64_t T, int N, int M, const double *logpi, const double *logA,
                              const double *logBsm, const void *obs, int obs_dtype, const int32_t *lengths,
                              void *backptr, size_t backptr_bytes, void *path, int path_dtype, double *logp,
                              khmm_stream_t stream) {
    KHMM_REQUIRE(B >= 0 && T >= 1 && N >= 1 && M >= 1, "viterbi: bad shape B=%lld T=%lld N=%d M=%d", (long long)B,
                 (long long)T, N, M);
    if (N > 1024) {
        set_error("N=%d > 1024 is not supported", N);
        return KHMM_ERR_UNSUPPORTED;
    }
    if (B == 0) return KHMM_OK;
    KHMM_REQUIRE(logpi && logA && logBsm && obs && path && logp && backptr, "viterbi: NULL pointer argument");
    KHMM_REQUIRE(obs_dtype >= KHMM_U8 && obs_dtype <= KHMM_I64, "viterbi: obs dtype code %d is not an integer type",
                 obs_dtype);
    KHMM_REQUIRE(path_dtype >= KHMM_U8 && path_dtype <= KHMM_I64, "viterbi: path dtype code %d is not an integer type",
                 path_dtype);
    KHMM_REQUIRE(backptr_bytes >= khmm_viterbi_workspace_bytes(B, T, N), "viterbi: workspace too small (%zu < %zu)",
                 backptr_bytes, khmm_viterbi_workspace_bytes(B, T, N));
    if (B == 0) return KHMM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 256)
        return run_viterbi<uint8_t>(B, T, N, M, logpi, logA, logBsm, obs, obs_dtype, lengths, backptr, path,
                                    path_dtype, logp, st);
    return run_viterbi<uint16_t>(B, T, N, M, logpi, logA, logBsm, obs, obs_dtype, lengths, backptr, path, path_dtype,
                                 logp, st);
}

}  // extern "C"
