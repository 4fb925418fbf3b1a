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
#include <stdlib.h>

namespace khmm {

constexpr int kVitSeq = 4;

template <typename PSI, bool A_SMEM, int S>
__global__ void __launch_bounds__(1024)
viterbi_generic_kernel(int64_t B, int64_t T, int N, int M, const double *__restrict__ logpi,
                       const double *__restrict__ logA, const double *__restrict__ logBsm, const void *__restrict__ obs,
                       int obs_dtype, const int32_t *__restrict__ lengths, PSI *__restrict__ psi,
                       int32_t *__restrict__ last_state, double *__restrict__ logp) {
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
    // M == 0: 1-D Gaussian emissions (extension, khmm_viterbi_gauss_f64); logBsm is then [3][N] = mean, sd, log-constant
    const bool gauss = M == 0;
    const double g_mean = (gauss && active) ? logBsm[j] : 0.0, g_sd = (gauss && active) ? logBsm[N + j] : 1.0,
                 g_logc = (gauss && active) ? logBsm[2 * N + j] : 0.0;
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
        // log b_j(o_t) is requested one step ahead (symbol -> table row: two dependent global loads)
        auto log_emission = [&](int s, int tt) -> double {
            if (!(active && tt < len[s])) return 0.0;
            if (gauss) {
                // z = (x - mean)/sd; log pdf = (-0.5 z) z + (-log sd - log sqrt(2 pi)): one rounding per operation, in the
                // order the oracle's NumPy expression evaluates (no contraction into an FMA)
                const double x = load_real<double>(obs, obs_dtype, (b0 + s) * T + tt);
                const double z = __ddiv_rn(__dsub_rn(x, g_mean), g_sd);
                return __dadd_rn(__dmul_rn(__dmul_rn(-0.5, z), z), g_logc);
            }
            int sym = load_symbol(obs, obs_dtype, (b0 + s) * T + tt);
            sym = sym < 0 ? 0 : (sym >= M ? M - 1 : sym);
            return __ldg(logBsm + (int64_t)sym * N + j);
        };
        double e_cur[S];
#pragma unroll
        for (int s = 0; s < S; s++) e_cur[s] = log_emission(s, 0);
        int cur = 0;
        for (int t = 0; t < tmax; t++) {
            const double *prev = rows + cur * NT * S;
            double *next = rows + (cur ^ 1) * NT * S;
            double best[S];
            int arg[S];
            double e_nxt[S];
#pragma unroll
            for (int s = 0; s < S; s++) e_nxt[s] = log_emission(s, t + 1);
            if (t == 0) {
#pragma unroll
                for (int s = 0; s < S; s++) { best[s] = lpj; arg[s] = 0; }
            } else if (active) {
                auto load_row = [&](int i, double (&p)[S]) {       // the S values delta_{t-1}[i] of the CTA's sequences
                    if constexpr (S == 4) {
                        const double2 p0 = *reinterpret_cast<const double2 *>(prev + i * S);
                        const double2 p1 = *reinterpret_cast<const double2 *>(prev + i * S + 2);
                        p[0] = p0.x; p[1] = p0.y; p[2] = p1.x; p[3] = p1.y;
                    } else {
#pragma unroll
                        for (int s = 0; s < S; s++) p[s] = prev[i * S + s];
                    }
                };
                {
                    const double a0 = Am[j];
                    double p[S];
                    load_row(0, p);
#pragma unroll
                    for (int s = 0; s < S; s++) { best[s] = p[s] + a0; arg[s] = 0; }
                }
#pragma unroll 4
                for (int i = 1; i < N; i++) {
                    const double aij = Am[i * N + j];
                    double p[S];
                    load_row(i, p);
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        const double v = p[s] + aij;
                        if (v > best[s]) { best[s] = v; arg[s] = i; }
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < S; s++) {
                double d = 0.0;
                if (active && t < len[s]) {
                    int64_t b = b0 + s;
                    d = best[s] + e_cur[s];
                    if (t > 0) psi[(b * T + t) * N + j] = (PSI)arg[s];
                }
                next[j * S + s] = d;
                e_cur[s] = e_nxt[s];
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

// ------------------------------------------------------------------ float32-screened kernel (N = 32, 64)
// The register-tile kernel spends five instructions per transition (DADD, DSETP, two FSEL, SEL) on the half-rate
// FP64 / ALU pipes.  Almost all of that work decides something a float32 comparison already decides: which source
// state wins.  This kernel finds the winner in float32 and evaluates ONLY the winner in float64:
//
//   x_i  = delta[i] - R,  R = fl32(max_i delta[i])   (float64 subtraction, once per source and step; <= 2^-24 |R|)
//   s_ij = fl32( fl32(x_i) + fl32(logA[i][j]) ) (screen value of the candidate  v_ij = fl64(delta[i] + logA[i][j]))
//   p_gj = max of s_ij over the four sources of group g,  m_j = max_g p_gj
//                                               (packed FADD2 + 5/8 max instruction per transition)
//   the GROUPS with  p_gj > m_j - D_j  are counted and their indices summed with two FMA-pipe instructions per
//   group ( c = fma.sat(p, 2^100, -thr 2^100) in {0, 1};  acc += c (g + 256) ).
//
// |s_ij - (v_ij - max delta)| <= 2^-23 (|s_ij| + 2 max(0, max logA)) (1 + 2^-22) + 2^-52 (|max delta| + |v_ij|): three
// float32 roundings of quantities bounded by |x_i| + |logA_ij|, plus the float64 roundings of x_i and v_ij.  With
// D_j = 2^-20 (|m_j| + 2 max(0, max logA)) + 2^-48 |max delta| + 2^-60 -- four times the sum of the two error bounds --
// every source that attains the float64 maximum passes the filter, and so does its group.  If exactly ONE group
// passes, all maximisers are among its four sources: those four are evaluated in float64 in index order with the
// reference's strict compare, so delta_t[j] and psi_t[j] are bit-identical to abstract_hmm.py:227-230.  Otherwise
// (a near-tie between groups, -inf rows, a threshold too close to zero) the lane runs the reference's float64
// loop over all sources with the first-index rule.  A warp that needed the float64 loop three steps
// in a row stops screening and re-probes after 16 steps, so tie-heavy models (left-to-right, uniform rows) cost
// ~6 % extra instead of a wasted screen per step.
//
// Mapping: one WARP per sequence (no CTA barriers in the time loop), 16 warps per CTA, lane l owns targets l, l+32;
// fl32(logA) in shared memory grouped by four sources (one conflict-free LDS.128 per group and target), the float64
// logA and the delta rows in shared memory for the exact evaluations.
// fl32(logA) lives in REGISTERS for N <= 64 (8 warps per CTA; 9+ warps leave 168 registers per thread and spill:
// 320-390 ms instead of 187 ms at c3) and in shared memory, grouped by four sources, for N = 128 (the same layout at
// N = 64 measured 224-230 ms against 187 ms for the register copy and 349 ms for the float64 kernel).  8 warps per
// CTA in both cases: at N = 128 sixteen warps leave 128 registers and spill (48.9 ms at B = 16384, T = 512, no
// better than the float64 register-tile kernel's 47.7 ms); eight warps with 255 registers run 24.6 ms.  A hybrid at
// N = 64 (one target per lane in registers, the other in shared memory; 10 / 12 / 16 warps) measured 224 / 198 / 230 ms.
#ifndef KHMM_VIT_JR64
#define KHMM_VIT_JR64 2        /* N = 64: targets per lane whose fl32(logA) column sits in registers (0..2) */
#endif
#ifndef KHMM_VIT_WARPS64
#define KHMM_VIT_WARPS64 8
#endif
template <int NN>
struct VitScreen {
    static constexpr int JT = NN / 32;
    static constexpr int JR = NN < 64 ? JT : (NN == 64 ? KHMM_VIT_JR64 : 0);   // targets per lane with a register copy
    static constexpr int WARPS = NN == 64 ? KHMM_VIT_WARPS64 : 8;
    static constexpr size_t smem_bytes() {
        return sizeof(double) * (NN * NN + WARPS * NN) + sizeof(float) * WARPS * NN + (JR == JT ? 0 : sizeof(float) * NN * NN);
    }
};
template <int NN, typename PSI, bool GAUSS>
__global__ void __launch_bounds__(32 * VitScreen<NN>::WARPS, 1)
viterbi_screen_kernel(int64_t B, int64_t T, int M, const double *__restrict__ logpi, const double *__restrict__ logA,
                      const double *__restrict__ logBsm, const void *__restrict__ obs, int obs_dtype,
                      const int32_t *__restrict__ lengths, PSI *__restrict__ psi, int32_t *__restrict__ last_state,
                      double *__restrict__ logp) {
    constexpr int JT = NN / 32, WARPS = VitScreen<NN>::WARPS, JR = VitScreen<NN>::JR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *A64 = reinterpret_cast<double *>(smem_raw);                 // [NN][NN]
    double *drow = A64 + NN * NN;                                       // [WARPS][NN]  delta_{t-1}, float64
    float *dlrow = reinterpret_cast<float *>(drow + WARPS * NN);        // [WARPS][NN]  fl32(delta - R)
    float4 *A32g = reinterpret_cast<float4 *>(dlrow + WARPS * NN);      // [NN/4][NN]   fl32(logA[4g .. 4g+3][j])
    __shared__ float apos_red[WARPS];
    float ap = 0.f;                                                     // max(0, largest finite logA), rounded up
    for (int k = threadIdx.x; k < NN * NN; k += blockDim.x) {
        const double v = logA[k];
        A64[k] = v;
        if (v > 0.0 && v < 1e300) ap = fmaxf(ap, __double2float_ru(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ap = fmaxf(ap, __shfl_xor_sync(0xffffffffu, ap, o));
    if ((threadIdx.x & 31) == 0) apos_red[threadIdx.x >> 5] = ap;
    __syncthreads();
    ap = apos_red[0];
#pragma unroll
    for (int k = 1; k < WARPS; k++) ap = fmaxf(ap, apos_red[k]);
    const float apos2 = 2.f * ap;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    float4 a4r[JR ? JR : 1][JR ? NN / 4 : 1];      // register copy: a4r[q][g] = fl32(logA[4g .. 4g+3][l + 32 q]), q < JR
#pragma unroll
    for (int q = 0; q < JR; q++)
#pragma unroll
        for (int g = 0; g < NN / 4; g++) {
            const int j = l + 32 * q;
            a4r[JR ? q : 0][JR ? g : 0] = make_float4((float)A64[(4 * g) * NN + j], (float)A64[(4 * g + 1) * NN + j],
                                                      (float)A64[(4 * g + 2) * NN + j], (float)A64[(4 * g + 3) * NN + j]);
        }
    if (JR < JT) {
        for (int k = threadIdx.x; k < (NN / 4) * NN; k += blockDim.x) {
            const int g = k / NN, j = k % NN;
            A32g[k] = make_float4((float)A64[(4 * g) * NN + j], (float)A64[(4 * g + 1) * NN + j], (float)A64[(4 * g + 2) * NN + j],
                                  (float)A64[(4 * g + 3) * NN + j]);
        }
        __syncthreads();
    }
    double lp[JT];
#pragma unroll
    for (int q = 0; q < JT; q++) lp[q] = logpi[l + 32 * q];
    double *dw = drow + w * NN;
    float *dlw = dlrow + w * NN;
    const float BIG = 0x1p100f;
    const float FINF = __int_as_float(0x7f800000);

    for (int64_t b = (int64_t)blockIdx.x * WARPS + w; b < B; b += (int64_t)gridDim.x * WARPS) {
        const int len = lengths ? lengths[b] : (int)T;
        if (len <= 0) continue;
        // observation token of step tt (clamped into the sequence): the symbol, or (GAUSS) the real observation; and the
        // log-emissions of this lane's targets for a token.  GAUSS (extension, khmm_viterbi_gauss_f64): logBsm holds
        // [3][NN] = mean, sd, log-constant and log b_j(x) = (-0.5 z) z + logc_j, z = (x - mean_j) / sd_j, one rounding per
        // operation in that order (no FMA contraction), as the generic kernel and the oracle evaluate it
        auto token_at = [&](int tt) -> double {
            const int64_t idx = b * T + (tt < len ? tt : len - 1);
            if (GAUSS) return load_real<double>(obs, obs_dtype, idx);
            int sy = load_symbol(obs, obs_dtype, idx);
            return (double)(sy < 0 ? 0 : (sy >= M ? M - 1 : sy));
        };
        auto log_emissions = [&](double tok, double (&eo)[JT]) {
#pragma unroll
            for (int q = 0; q < JT; q++) {
                const int j = l + 32 * q;
                if (GAUSS) {
                    const double z = __ddiv_rn(__dsub_rn(tok, __ldg(logBsm + j)), __ldg(logBsm + NN + j));
                    eo[q] = __dadd_rn(__dmul_rn(__dmul_rn(-0.5, z), z), __ldg(logBsm + 2 * NN + j));
                } else {
                    eo[q] = __ldg(logBsm + (int64_t)(int)tok * NN + j);
                }
            }
        };
        double d[JT];
        {
            double e0[JT];
            log_emissions(token_at(0), e0);
#pragma unroll
            for (int q = 0; q < JT; q++) d[q] = lp[q] + e0[q];
        }
        int exact_until = 0, streak = 0;       // steps < exact_until skip the screen
        // the token of step t+2 and the emission row of step t+1 are requested during step t (both are L2 latencies)
        double e_next[JT];
        log_emissions(token_at(1), e_next);
        double tok_next = token_at(2);
        for (int t = 1; t < len; t++) {
            double e[JT];
#pragma unroll
            for (int q = 0; q < JT; q++) e[q] = e_next[q];
            log_emissions(tok_next, e_next);
            tok_next = token_at(t + 2);
            // reference level R = fl32(max of the previous row): one integer warp reduction on the order-preserving
            // integer image of the float32 values (the screen only needs delta - R to be small near the top)
            float lf = (float)d[0];
#pragma unroll
            for (int q = 1; q < JT; q++) lf = fmaxf(lf, (float)d[q]);
            int key = __float_as_int(lf);
            key = key >= 0 ? key : key ^ 0x7fffffff;
            key = __reduce_max_sync(0xffffffffu, key);
            key = key >= 0 ? key : key ^ 0x7fffffff;
            const float Rf = __int_as_float(key);
            const double dmax = (double)Rf;
            const bool screen = t >= exact_until && Rf > -FINF && Rf < FINF;
            __syncwarp();                      // the previous step's readers are done with dw / dlw
#pragma unroll
            for (int q = 0; q < JT; q++) {
                dw[l + 32 * q] = d[q];
                if (screen) dlw[l + 32 * q] = (float)(d[q] - dmax);
            }
            __syncwarp();
            double best[JT];
            int arg[JT];
            bool done_q[JT];
#pragma unroll
            for (int q = 0; q < JT; q++) { done_q[q] = false; best[q] = 0.0; arg[q] = 0; }
            if (screen) {
                // max delta - R <= 2^-24 |R|: the x_i may be that much above zero, which adds twice that to the bound on
                // |x_i| + |logA_ij| in terms of |s_ij| (see the header)
                const float apx = fmaf(fabsf(Rf), 0x1p-22f, apos2);
                const float dmx_term = fmaf(fabsf(Rf), 0x1p-48f, 0x1p-60f) * 1.0000002f;
#pragma unroll
                for (int q = 0; q < JT; q++) {
                    // group maxima p[g] = max of the four screen values of sources 4g .. 4g+3 (two packed adds, two max
                    // instructions per group); only the groups are thresholded, the winning group is resolved in float64
                    float p[NN / 4];
                    float m0 = -FINF, m1 = m0;
#pragma unroll
                    for (int g = 0; g < NN / 4; g++) {
                        const float4 x4 = *reinterpret_cast<const float4 *>(dlw + 4 * g);
                        const float4 a4 = q < JR ? a4r[q < JR ? q : 0][JR ? g : 0] : A32g[g * NN + l + 32 * q];
                        const float2 s01 = __fadd2_rn(make_float2(x4.x, x4.y), make_float2(a4.x, a4.y));
                        const float2 s23 = __fadd2_rn(make_float2(x4.z, x4.w), make_float2(a4.z, a4.w));
                        float pg = s01.x;
                        asm("max.f32 %0, %0, %1, %2;" : "+f"(pg) : "f"(s01.y), "f"(s23.x));
                        p[g] = fmaxf(pg, s23.y);
                    }
#pragma unroll
                    for (int g = 0; g < NN / 4; g += 4) {
                        asm("max.f32 %0, %0, %1, %2;" : "+f"(m0) : "f"(p[g]), "f"(p[g + 1]));
                        asm("max.f32 %0, %0, %1, %2;" : "+f"(m1) : "f"(p[g + 2]), "f"(p[g + 3]));
                    }
                    const float m = fmaxf(m0, m1);
                    const float dlt = fmaf(fabsf(m) + apx, 0x1p-20f, dmx_term);
                    const float thr = m - dlt;
                    const float nthrB = -thr * BIG;
                    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                    for (int g = 0; g < NN / 4; g += 2) {
                        float c0, c1;
                        asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(c0) : "f"(p[g]), "f"(BIG), "f"(nthrB));
                        asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(c1) : "f"(p[g + 1]), "f"(BIG), "f"(nthrB));
                        acc0 = fmaf(c0, (float)(g + 256), acc0);
                        acc1 = fmaf(c1, (float)(g + 257), acc1);
                    }
                    const float acc = acc0 + acc1;
                    // exactly one group above the threshold (and a threshold far enough from zero for c to be 0 or 1):
                    // every source that attains the float64 maximum is in that group -- evaluate its four sources exactly
                    if (acc >= 256.f && acc < 512.f && fabsf(thr) >= 0x1p-60f) {
                        const int k = ((int)acc - 256) * 4;
                        const int j = l + 32 * q;
                        double bb = dw[k] + A64[k * NN + j];
                        int ba = k;
#pragma unroll
                        for (int e2 = 1; e2 < 4; e2++) {
                            const double v = dw[k + e2] + A64[(k + e2) * NN + j];
                            if (v > bb) { bb = v; ba = k + e2; }
                        }
                        arg[q] = ba;
                        best[q] = bb;
                        done_q[q] = true;
                    }
                }
            }
            bool all_done = true;
#pragma unroll
            for (int q = 0; q < JT; q++) all_done = all_done && done_q[q];
            if (!__all_sync(0xffffffffu, all_done)) {
                // the reference's loop, float64, first index wins ties (abstract_hmm.py:227-230)
#pragma unroll
                for (int q = 0; q < JT; q++) {
                    if (!done_q[q]) {
                        double bb = dw[0] + A64[l + 32 * q];
                        int ba = 0;
#pragma unroll 4
                        for (int i = 1; i < NN; i++) {
                            const double v = dw[i] + A64[i * NN + l + 32 * q];
                            if (v > bb) { bb = v; ba = i; }
                        }
                        best[q] = bb;
                        arg[q] = ba;
                    }
                }
                // three steps in a row needed the float64 loop: stop screening for 16 steps
                if (screen && ++streak >= 3) { exact_until = t + 16; streak = 0; }
            } else {
                streak = 0;
            }
#pragma unroll
            for (int q = 0; q < JT; q++) {
                psi[(b * T + t) * NN + l + 32 * q] = (PSI)arg[q];
                d[q] = best[q] + e[q];
            }
        }
        // termination: first-index argmax over the final row (abstract_hmm.py:233,235)
        double m = d[0];
        int am = l;
#pragma unroll
        for (int q = 1; q < JT; q++)
            if (d[q] > m) { m = d[q]; am = l + 32 * q; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double om = __shfl_xor_sync(0xffffffffu, m, o);
            const int oa = __shfl_xor_sync(0xffffffffu, am, o);
            if (om > m || (om == m && oa < am)) { m = om; am = oa; }
        }
        if (l == 0) {
            last_state[b] = am;
            logp[b] = m;
        }
    }
}

template <int NN, typename PSI, bool GAUSS>
static int launch_viterbi_screen(int64_t B, int64_t T, int M, const double *logpi, const double *logA, const double *logBsm,
                                 const void *obs, int obs_dtype, const int32_t *lengths, PSI *psi, int32_t *last,
                                 double *logp, cudaStream_t st) {
    auto kern = viterbi_screen_kernel<NN, PSI, GAUSS>;
    constexpr int kVitScreenWarps = VitScreen<NN>::WARPS;
    const size_t smem = VitScreen<NN>::smem_bytes();
    if (smem > 48 * 1024) KHMM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    KHMM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * kVitScreenWarps, smem));
    if (occ < 1) occ = 1;
    const int64_t nblk = (B + kVitScreenWarps - 1) / kVitScreenWarps, cap = (int64_t)sm_count() * occ;   // persistent: one resident wave of warps
    kern<<<(int)(nblk < cap ? nblk : cap), 32 * kVitScreenWarps, smem, st>>>(B, T, M, logpi, logA, logBsm, obs, obs_dtype, lengths, psi, last, logp);
    KHMM_LAUNCH_CHECK("viterbi_screen_kernel");
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

// Small batches: a WARP per sequence stages the backpointer rows of 32 steps in shared memory (they are contiguous:
// psi[b][t][:]) with coalesced loads and one lane chases through the staged rows -- one DRAM latency per 32 steps
// instead of one per step.  The thread-per-sequence kernel above is bound by that latency whatever the batch size
// (T x ~0.55 us: 2.2 ms at T = 4096, a fifth of a lone T = 1000 sequence's three calls); it stays the kernel of large
// batches, where reading every row (N bytes per step instead of one 32-byte sector) would cost more than the latency.
constexpr int kBtChunk = 32, kBtWarps = 4;
template <typename OUT>
__global__ void __launch_bounds__(32 * kBtWarps)
viterbi_backtrace_staged_kernel(int64_t B, int64_t T, int N, const int32_t *__restrict__ lengths,
                                const uint8_t *__restrict__ psi, const int32_t *__restrict__ last_state,
                                OUT *__restrict__ path) {
    extern __shared__ __align__(16) unsigned char bt_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * kBtWarps + warp;
    if (b >= B) return;
    const int rows_bytes = (kBtChunk * N + 15) & ~15;
    unsigned char *rows = bt_smem + (size_t)warp * (rows_bytes + kBtChunk * sizeof(int));
    int *out = reinterpret_cast<int *>(rows + rows_bytes);
    const int len = lengths ? lengths[b] : (int)T;
    int st = last_state[b];
    OUT *row = path + b * T;
    const uint8_t *prow = psi + b * T * N;
    if (lane == 0) row[len - 1] = (OUT)st;
    const bool vec = (N % 16 == 0) && ((reinterpret_cast<uintptr_t>(psi) & 15) == 0);
    for (int t_hi = len - 2; t_hi >= 0; t_hi -= kBtChunk) {
        const int t_lo = max(t_hi - kBtChunk + 1, 0), n = t_hi - t_lo + 1;
        const uint8_t *src = prow + (int64_t)(t_lo + 1) * N;          // rows t_lo+1 .. t_hi+1, contiguous
        const int nbytes = n * N;
        if (vec) {
            for (int i = lane * 16; i < nbytes; i += 32 * 16)
                *reinterpret_cast<uint4 *>(rows + i) = *reinterpret_cast<const uint4 *>(src + i);
        } else {
            for (int i = lane; i < nbytes; i += 32) rows[i] = src[i];
        }
        __syncwarp();
        if (lane == 0) {
            for (int j = n - 1; j >= 0; j--) {                        // step t = t_lo + j reads row t + 1 = staged row j
                st = (int)rows[j * N + st];
                out[j] = st;
            }
        }
        st = __shfl_sync(0xffffffffu, st, 0);
        __syncwarp();
        if (lane < n) row[t_lo + lane] = (OUT)out[lane];
        __syncwarp();
    }
}

template <typename PSI>
static int run_viterbi(int64_t B, int64_t T, int N, int M, const double *logpi, const double *logA,
                       const double *logBsm, const void *obs, int obs_dtype, const int32_t *lengths, void *backptr,
                       void *path, int path_dtype, double *logp, cudaStream_t st) {
    // small batches: one sequence per CTA (a lone warp then issues a quarter of the instructions per step)
    const int S = (B + kVitSeq - 1) / kVitSeq < sm_count() ? 1 : kVitSeq;
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
    const bool use_screen = dev_switch(SW_VIT_SCREEN) != 0;      // development / cross-check switch (khmm_debug_switch)
    // N <= 64: from one sequence up (T = 4096: 3.2 ms against 6.3 ms on the generic kernel at N = 64, 2.2 against 4.2 ms
    // at N = 32); N = 128: from 64 sequences (every CTA first copies 192 KB of logA into shared memory: 14.2 against
    // 10.7 ms at B = 8)
    if (sizeof(PSI) == 1 && B >= (N <= 64 ? 1 : 64) && use_screen && (N == 32 || N == 64 || N == 128)) {
        // float32-screened warp-per-sequence kernel (M == 0: Gaussian emissions, logBsm = [3][N] mean / sd / log-constant)
        if (M > 0) {
            if (N == 32) rc_tile = launch_viterbi_screen<32, PSI, false>(B, T, M, logpi, logA, logBsm, obs, obs_dtype, lengths, psi, last, logp, st);
            else if (N == 64) rc_tile = launch_viterbi_screen<64, PSI, false>(B, T, M, logpi, logA, logBsm, obs, obs_dtype, lengths, psi, last, logp, st);
            else rc_tile = launch_viterbi_screen<128, PSI, false>(B, T, M, logpi, logA, logBsm, obs, obs_dtype, lengths, psi, last, logp, st);
        } else {
            if (N == 32) rc_tile = launch_viterbi_screen<32, PSI, true>(B, T, M, logpi, logA, logBsm, obs, obs_dtype, lengths, psi, last, logp, st);
            else if (N == 64) rc_tile = launch_viterbi_screen<64, PSI, true>(B, T, M, logpi, logA, logBsm, obs, obs_dtype, lengths, psi, last, logp, st);
            else rc_tile = launch_viterbi_screen<128, PSI, true>(B, T, M, logpi, logA, logBsm, obs, obs_dtype, lengths, psi, last, logp, st);
        }
    } else if (M > 0 && sizeof(PSI) == 1 && B >= 64) {   // register-tile kernels (small batches stay on the generic kernel)
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
        auto kern = S == 1 ? (as ? viterbi_generic_kernel<PSI, true, 1> : viterbi_generic_kernel<PSI, false, 1>)
                           : (as ? viterbi_generic_kernel<PSI, true, kVitSeq> : viterbi_generic_kernel<PSI, false, kVitSeq>);
        if (smem > 48 * 1024)
            KHMM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, NT, smem, st>>>(B, T, N, M, logpi, logA, logBsm, obs, obs_dtype, lengths, psi, last, logp);
        KHMM_LAUNCH_CHECK("viterbi_generic_kernel");
    }
    if (sizeof(PSI) == 1 && B * (int64_t)N <= (1 << 21)) {
        // small batch: staged warp-per-sequence backtrace (one memory latency per 32 steps)
        const size_t sm = (size_t)kBtWarps * (((size_t)kBtChunk * N + 15) / 16 * 16 + kBtChunk * sizeof(int));
        const unsigned g = (unsigned)((B + kBtWarps - 1) / kBtWarps);
        const uint8_t *p8 = reinterpret_cast<const uint8_t *>(psi);
        switch (path_dtype) {
            case KHMM_U8: viterbi_backtrace_staged_kernel<uint8_t><<<g, 32 * kBtWarps, sm, st>>>(B, T, N, lengths, p8, last, (uint8_t *)path); break;
            case KHMM_U16: viterbi_backtrace_staged_kernel<uint16_t><<<g, 32 * kBtWarps, sm, st>>>(B, T, N, lengths, p8, last, (uint16_t *)path); break;
            case KHMM_I32: viterbi_backtrace_staged_kernel<int32_t><<<g, 32 * kBtWarps, sm, st>>>(B, T, N, lengths, p8, last, (int32_t *)path); break;
            default: viterbi_backtrace_staged_kernel<int64_t><<<g, 32 * kBtWarps, sm, st>>>(B, T, N, lengths, p8, last, (int64_t *)path); break;
        }
        KHMM_LAUNCH_CHECK("viterbi_backtrace_staged_kernel");
        return KHMM_OK;
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

int khmm_viterbi_gauss_f64(int64_t B, int64_t T, int N, const double *logpi, const double *logA,
                           const double *mean_sd_logc, const void *obs, int obs_dtype, const int32_t *lengths,
                           void *backptr, size_t backptr_bytes, void *path, int path_dtype, double *logp,
                           khmm_stream_t stream) {
    KHMM_REQUIRE(B >= 0 && T >= 1 && N >= 1, "viterbi_gauss: bad shape B=%lld T=%lld N=%d", (long long)B, (long long)T, N);
    if (N > 1024) {
        set_error("N=%d > 1024 is not supported", N);
        return KHMM_ERR_UNSUPPORTED;
    }
    if (B == 0) return KHMM_OK;
    KHMM_REQUIRE(logpi && logA && mean_sd_logc && obs && path && logp && backptr, "viterbi_gauss: NULL pointer argument");
    KHMM_REQUIRE(obs_dtype == KHMM_F32 || obs_dtype == KHMM_F64, "viterbi_gauss: obs dtype code %d is not f32/f64", obs_dtype);
    KHMM_REQUIRE(path_dtype >= KHMM_U8 && path_dtype <= KHMM_I64, "viterbi_gauss: path dtype code %d is not an integer type",
                 path_dtype);
    KHMM_REQUIRE(backptr_bytes >= khmm_viterbi_workspace_bytes(B, T, N), "viterbi_gauss: workspace too small (%zu < %zu)",
                 backptr_bytes, khmm_viterbi_workspace_bytes(B, T, N));
    cudaStream_t st = (cudaStream_t)stream;
    // M = 0 selects the Gaussian emission evaluation of the generic kernel
    if (N <= 256)
        return run_viterbi<uint8_t>(B, T, N, 0, logpi, logA, mean_sd_logc, obs, obs_dtype, lengths, backptr, path, path_dtype,
                                    logp, st);
    return run_viterbi<uint16_t>(B, T, N, 0, logpi, logA, mean_sd_logc, obs, obs_dtype, lengths, backptr, path, path_dtype,
                                 logp, st);
}

}  // extern "C"
