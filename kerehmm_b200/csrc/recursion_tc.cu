// Tensor-core recursion for large state spaces (N a multiple of 128, 128 <= N <= 1024): the forward step
// alpha_{t-1} . A (abstract_hmm.py:266-270) and the backward step A . (b . beta_{t+1}) (:339-349) of
// a whole batch are one (B x N) x (N x N) contraction per time step, run on tcgen05 (kind::tf32 or kind::f16 with
// bf16 operands, fp32 accumulators in TMEM, operands staged by TMA with the 128-byte swizzle) with an epilogue that
// fuses the emission evaluation, the scaling and the row sums:
//
//     W_t = (W_{t-1} M) . b_t . r_{t-1},      r_{t-1} = 1 / sum_j W_{t-1}(j),
//
// M = A with reversed time gives the (self-scaled) backward recursion in the form w_t = b_t . beta_t,
// exactly as in smalln.cu.  W_t is the reference's *unnormalised* alpha_t (its row sum is the scale
// factor c_t), so normalisation needs no second pass over a row that is split across CTAs: every
// column tile writes its partial row sums and the next step's epilogue adds them up.  Forward and
// backward run together, log P accumulates in float64 from the c_t.
//
// Two drivers of that step:
//   * tc_recursion2_kernel -- ALL T-1 steps in one cooperative launch, jobs chained through per-chain completion
//     counters (the product path; see the comment above it and DESIGN.md 3.4);
//   * tc_step_kernel -- one launch per time step (devices without cooperative launch, KHMM_RC_MODE=1): warps 0-3
//     epilogue (thread = TMEM lane = sequence), warp 4 TMA producer, warp 5 TMEM allocator + single-thread MMA issuer,
//     4-stage smem ring.
//
// Precision: tf32 multiplies (10-bit mantissa; operands are ROUNDED to tf32 when written, the tensor core would
// truncate) or bf16 multiplies, fp32 accumulation.  Measured against the float64 oracle at N = 256..1024 the
// log-likelihood agrees to ~1e-6 relative with tf32 (tests/test_gpu_parity.py states the bounds); alpha itself is only
// good to ~1e-3, which is why the float32 SIMT kernels remain the parity path for alpha/beta outputs.
#include "common.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

#include "tcgen05.cuh"

namespace khmm {

using namespace tc;

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32;   // TC_BN: psum / workspace granularity (column tile of 128)
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;       // one 128 x 32 fp32 operand tile
constexpr int TC_STAGES = 4;                           // ring depth (BN = 128: 32 KB per stage, BN = 256: 48 KB)

struct __align__(8) TcBars {
    uint64_t full[TC_STAGES], empty[TC_STAGES], tmem_full;
    uint32_t tmem_base;
};

struct TcDir {              // per direction (0 forward, 1 backward)
    float *W;               // [B][2][N] ping-pong rows (unnormalised)
    float *psum;            // [B][2][NT] partial row sums per column tile
    double *logacc;         // [B] running sum of log c
};

struct TcEmis {
    int gauss;              // 1: 1-D Gaussian (mean, sd), 0: discrete table
    const float *p0, *p1;   // gauss: mean, sd ; discrete: Bsm [M][N]
    const void *obs;
    int obs_dtype, M;
    // ragged batches (whole-recursion kernel only): 1 <= lengths[b] <= T, rows padded to T; NULL = equal lengths.  Both
    // recursions run on through the padding with the emission factor set to 1 there (every value stays finite whatever
    // the padding holds); the forward pass leaves the padding's scale factors out of log P; the backward pass holds
    // w = 1 through the padding and restarts exactly at the row's last step, w_{len-1} = b(o_{len-1}).
    const int32_t *lengths;
};

// Round to nearest tf32 (10-bit mantissa).  The tensor core TRUNCATES fp32 operands to tf32, which
// biases every product low (measured: log-likelihood off by -6.6e-4 per step); operands that are
// already tf32-representable pass through exactly, so rounding them here makes the error unbiased.
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__global__ void tc_to_bf16_kernel(const float *__restrict__ in, __nv_bfloat16 *__restrict__ out, int n) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) out[k] = __float2bfloat16_rn(in[k]);
}

__global__ void tc_round_matrix_kernel(const float *__restrict__ in, float *__restrict__ out, int n) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) out[k] = round_tf32(in[k]);
}

template <bool GAUSS, int BN>
__device__ __forceinline__ float tc_emission(const TcEmis &em, const float *sp, int c, float x, const float *brow) {
    if (GAUSS) {
        float d = x - sp[c];
        float a = fmaf(d * d, sp[BN + c], sp[2 * BN + c]);
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
        return y;
    }
    return __ldg(brow + c);
}

// One time step for both directions.  step >= 1; forward handles time t = step, backward t = T-1-step.
template <bool GAUSS, int BN>
__global__ void __launch_bounds__(192, 1)
tc_step_kernel(const __grid_constant__ CUtensorMap mapXf, const __grid_constant__ CUtensorMap mapXb,
               const __grid_constant__ CUtensorMap mapMf, const __grid_constant__ CUtensorMap mapMb, TcDir df, TcDir db,
               TcEmis em, int64_t B, int64_t T, int N, int step) {
    constexpr int STAGE_BYTES = TC_TILE_BYTES + BN * TC_BK * 4;   // W tile + BN x 32 matrix tile
    constexpr int SUB = BN / 128;                                 // 128-column psum sub-tiles per CTA
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    TcBars *bars = (TcBars *)(tiles + TC_STAGES * STAGE_BYTES);
    float *sp = (float *)(bars + 1);  // [3][128] emission constants of this column tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.z;
    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
    const int nkb = N / TC_BK, NT = N / 128;
    const CUtensorMap *mapX = dir ? &mapXb : &mapXf;
    const CUtensorMap *mapM = dir ? &mapMb : &mapMf;
    const TcDir d = dir ? db : df;
    const int src = (step - 1) & 1, dst = step & 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->tmem_full, 1);
        fence_barrier_init();
    }
    if (GAUSS) {
        for (int c = threadIdx.x; c < BN; c += blockDim.x) {
            int j = n0 + c;
            float sd = em.p1[j];
            sp[c] = em.p0[j];
            sp[BN + c] = -0.5f * 1.4426950408889634f / (sd * sd);
            sp[2 * BN + c] = -log2f(sd * 2.5066282746310002f);
        }
    }
    if (warp == 5) tmem_alloc(&bars->tmem_base, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(mapX);
            tma_prefetch_desc(mapM);
            for (int kb = 0; kb < nkb; kb++) {
                int s = kb % TC_STAGES;
                uint32_t ph = (kb / TC_STAGES) & 1;
                mbar_wait(&bars->empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&bars->full[s], STAGE_BYTES);
                // W rows live in a [B][2][N] buffer viewed as a 2-D matrix [2B][N]: row = 2*b + slot
                tma_load_2d(tiles + s * STAGE_BYTES, mapX, &bars->full[s], kb * TC_BK, m0);
                tma_load_2d(tiles + s * STAGE_BYTES + TC_TILE_BYTES, mapM, &bars->full[s], kb * TC_BK, n0);
            }
        }
    } else if (warp == 5) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_tf32(TC_BM, BN);
            for (int kb = 0; kb < nkb; kb++) {
                int s = kb % TC_STAGES;
                uint32_t ph = (kb / TC_STAGES) & 1;
                mbar_wait(&bars->full[s], ph);
                tc_fence_after();
                uint64_t da = make_kmajor_sw128_desc(tiles + s * STAGE_BYTES);
                uint64_t dbb = make_kmajor_sw128_desc(tiles + s * STAGE_BYTES + TC_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; k++)
                    mma_tf32_ss(tmem, desc_advance(da, k * 32), desc_advance(dbb, k * 32), idesc, (kb | k) != 0);
                tc_commit(&bars->empty[s]);
            }
            tc_commit(&bars->tmem_full);
        }
    } else {
        // epilogue: thread = sequence row.  Everything that does not need the accumulator is done
        // while the MMAs run.
        const int64_t b = m0 + warp * 32 + lane;
        const bool live = b < B;
        const int64_t t = dir ? (T - 1 - step) : step;
        float r = 0.f, x = 0.f;
        const float *brow = nullptr;
        if (live) {
            const float *ps = d.psum + (b * 2 + src) * NT;
            float c = 0.f;
            for (int k = 0; k < NT; k++) c += ps[k];
            r = 1.0f / c;
            if (blockIdx.y == 0) d.logacc[b] += log((double)c);
            if (GAUSS) {
                x = load_real<float>(em.obs, em.obs_dtype, b * T + t);
            } else {
                int sym = load_symbol(em.obs, em.obs_dtype, b * T + t);
                sym = sym < 0 ? 0 : (sym >= em.M ? em.M - 1 : sym);
                brow = em.p0 + (int64_t)sym * N + n0;
            }
        }
        mbar_wait(&bars->tmem_full, 0);
        tc_fence_after();
        float rowsum = 0.f;
        float *out = d.W + (b * 2 + dst) * N + n0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (SUB > 1 && c0 == 128) {
                if (live) d.psum[(b * 2 + dst) * NT + blockIdx.y * SUB] = rowsum;
                rowsum = 0.f;
            }
            uint32_t acc[32];
            tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, acc);
            tmem_ld_wait();
            if (live) {
                float v[32];
#pragma unroll
                for (int q = 0; q < 32; q++) {
                    v[q] = __uint_as_float(acc[q]) * r * tc_emission<GAUSS, BN>(em, sp, c0 + q, x, brow);
                    rowsum += v[q];
                }
#pragma unroll
                for (int q = 0; q < 8; q++)
                    reinterpret_cast<float4 *>(out + c0)[q] = make_float4(round_tf32(v[4 * q]), round_tf32(v[4 * q + 1]),
                                                                          round_tf32(v[4 * q + 2]), round_tf32(v[4 * q + 3]));
            }
        }
        if (live) d.psum[(b * 2 + dst) * NT + blockIdx.y * SUB + SUB - 1] = rowsum;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, BN);
}

// helpers of the whole-recursion kernel
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {   // round to nearest even, lo in the low half
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_chain(const uint32_t *ctr, uint32_t need) {
    for (uint32_t spin = 0; ld_acquire_u32(ctr) < need; ++spin) {
        __nanosleep(64);
        if (spin > (1u << 26)) __trap();      // a scheduling bug traps instead of hanging the GPU
    }
}


// ------------------------------------------------------------------------------------------------------
// Whole-recursion kernel (second generation; the first one, tc_recursion_kernel, is in the history of this file:
// one epilogue warpgroup, lanes dealt to CTAs once).  ONE cooperative launch runs every time step: the work of a step
// is a set of independent CHAINS -- (row tile of 128 sequences, direction) -- each made of N/BN column-tile jobs that
// only depend on the same chain's jobs of the previous step; a job waits for its inputs on a per-chain completion
// counter in global memory (release by the producing epilogue, acquire by the consuming TMA thread) instead of on a
// kernel boundary, and only ever for jobs that precede it in every CTA's order, so the globally smallest unfinished
// job can always run (all CTAs are co-resident).  The clock64 timeline of the first generation (tools/recursion_tc_timeline.py)
// showed a job's MMAs taking 8.7k cycles (bf16) while its epilogue took 19k -- 9.5k of exposed L2 latency before the
// accumulator is even needed (serial row-sum loads, emission constants, a CTA barrier) and 9.4k of TMEM reads,
// emission evaluation and thread-per-row stores -- with ONE epilogue warpgroup serialising both; and 108 of the 148
// CTAs owned two lanes while 40 owned one.  Changes:
//   * two epilogue warpgroups (warps 0-3 / 4-7), one per TMEM accumulator: the epilogues of consecutive jobs overlap;
//   * a job's global order is g = (step-1)*nlanes + lane and walker w runs g = w, w+W, w+2W, ... across step
//     boundaries, so every walker gets the same number of jobs (chains drift apart freely, as before);
//   * lanes are chain-major (the column tiles of one chain are neighbours), which puts nlanes - ntile_n + 1 jobs
//     between a chain's step and its next step instead of nlanes - 3*mtiles: more slack for the dependency;
//   * Gaussian constants of all N states computed once per CTA; observation prefetched before the dependency
//     wait; the previous step's partial row sums fetched with independent vector loads; TMEM loads of the next
//     32-column chunk issued under the arithmetic of the current one; W_t staged in shared memory and written by TMA tensor stores;
//   * PAIR: a cluster of two CTAs (tcgen05 cta_group::2) runs a 256-sequence x 256-state job: each CTA stages its
//     own 128 W rows and HALF of the matrix tile, 512 KB instead of 768 KB of L2 -> SM traffic per 128 x 256 x 1024
//     job.  At c5 the single-CTA kernel asks the L2 for 256 x 1023 x 768 KB = 201 GB per call; the L2 delivers
//     ~6.3 KB/clk chip-wide, i.e. >= 16.8 ms, above the 10.9 ms the tensor pipe needs (bf16).
template <bool PAIR, int BN>
struct __align__(8) TcRBars {
    // operand ring: 16 KB W tile + (BN or BN/2) x 128 B matrix tile per stage; 144-160 KB, next to 32 KB of output staging
    // and the 12 KB constant table
    static constexpr int STAGES = PAIR ? 5 : (BN == 256 ? 3 : 5);
    uint64_t full[STAGES], empty[STAGES], tmem_full[4], tmem_empty[4];
    uint32_t tmem_base;
};
#ifndef KHMM_RC2
#define KHMM_RC2 1
#endif
#ifndef KHMM_RC2_PAIR
#define KHMM_RC2_PAIR 1
#endif
#ifndef KHMM_RC2_ORDER
#define KHMM_RC2_ORDER 0      /* 0: chain-major lanes, 1: row tile fastest (neighbouring walkers share a matrix tile) */
#endif

struct RcWalk {               // (step, lane) of the jobs w, w + stride, w + 2 stride, ... of the global order
    int64_t step;
    int ln, nlanes;
    __device__ __forceinline__ RcWalk(int start, int nl) : step(1), ln(start), nlanes(nl) { norm(); }
    __device__ __forceinline__ void norm() { while (ln >= nlanes) { ln -= nlanes; step++; } }
    __device__ __forceinline__ void advance(int by) { ln += by; norm(); }
};

__device__ __forceinline__ void mma2_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}

template <bool GAUSS, bool BF16, bool PAIR, int BN, bool RAGGED>
__device__ __forceinline__ void tc_recursion2_body(const CUtensorMap &mapXf0, const CUtensorMap &mapXf1, const CUtensorMap &mapXb0,
                                                   const CUtensorMap &mapXb1, const CUtensorMap &mapMf, const CUtensorMap &mapMb,
                                                   TcDir df, TcDir db, TcEmis em, int64_t B, int64_t T, int N, int mtiles,
                                                   int ndir, uint32_t *done, long long *trace) {
#define RC2_TRACE(ev, njv)                                                                              \
    do {                                                                                                \
        if (trace && blockIdx.x == 0 && (njv) >= 16 && (njv) < 24 && (threadIdx.x & 31) == 0)            \
            trace[((njv) - 16) * 16 + (ev)] = clock64();                                                 \
    } while (0)
    static_assert(BN == 256 || (BN == 128 && !PAIR), "column tiles: 256 (single CTAs or pairs) or 128 (single CTAs)");
    constexpr int STAGES = TcRBars<PAIR, BN>::STAGES;
    constexpr int MROWS = PAIR ? BN / 2 : BN;                       // matrix-tile rows staged by this CTA
    constexpr int HALF = BN / 2;                                    // columns per epilogue warp (column half h)
    constexpr int NACC = 512 / BN;                                  // accumulators in TMEM: the MMA thread may run NACC-1 jobs ahead
    constexpr int STAGE_BYTES = TC_TILE_BYTES + MROWS * 128;        // k-block rows are 128 bytes in both precisions
    constexpr int BKE = BF16 ? 64 : TC_BK;                          // elements per 128-byte k-block row
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *stage_out = tiles + STAGES * STAGE_BYTES;        // [2 warpgroups][128 rows][128 B], 128-byte swizzle
    TcRBars<PAIR, BN> *bars = (TcRBars<PAIR, BN> *)(stage_out + 2 * TC_TILE_BYTES);
    float *sp = (float *)(((uintptr_t)(bars + 1) + 15) & ~(uintptr_t)15);   // [3][N] Gaussian constants of all states
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    const int w = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // walker index
    const int nwalk = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int nkb = N / BKE, NTq = N / 64, ntile_n = N / BN;        // row sums are kept per 64 columns: [b][slot][N/64]
    const int mjobs = PAIR ? mtiles / 2 : mtiles;
    const int nlanes = mjobs * ntile_n * ndir;
    const int64_t njobs_total = (int64_t)nlanes * (T - 1);
    const uint32_t per_step = 2u * (uint32_t)ntile_n;                        // half-tile completions per chain and step

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int a = 0; a < NACC; a++) { mbar_init(&bars->tmem_full[a], 1); mbar_init(&bars->tmem_empty[a], PAIR ? 16 : 8); }
        fence_barrier_init();
    }
    if (GAUSS) {
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
            float sd = em.p1[j];
            sp[j] = -em.p0[j];
            sp[N + j] = -0.5f * 1.4426950408889634f / (sd * sd);
            sp[2 * N + j] = -log2f(sd * 2.5066282746310002f);
        }
    }
    if (warp == 9) { if (PAIR) tmem_alloc2(&bars->tmem_base, 512); else tmem_alloc(&bars->tmem_base, 512); }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    auto decode = [&](int ln, int &m0, int &n0, int &dir, int &chain) {
        int mtp, nt;
        if (KHMM_RC2_ORDER == 0) {
            nt = ln % ntile_n;
            const int cp = ln / ntile_n;
            dir = cp / mjobs;
            mtp = cp % mjobs;
        } else {
            mtp = ln % mjobs;
            const int rest = ln / mjobs;
            nt = rest % ntile_n;
            dir = rest / ntile_n;
        }
        const int mt = PAIR ? mtp * 2 + (int)rank : mtp;
        m0 = mt * TC_BM;
        n0 = nt * BN;
        chain = dir * mtiles + mt;
    };
    const int64_t myjobs = njobs_total > w ? (njobs_total - w + nwalk - 1) / nwalk : 0;   // jobs of this walker

    if (warp == 8) {
        if (elect_one()) {
            tma_prefetch_desc(&mapXf0); tma_prefetch_desc(&mapXf1); tma_prefetch_desc(&mapMf);
            if (ndir > 1) { tma_prefetch_desc(&mapXb0); tma_prefetch_desc(&mapXb1); tma_prefetch_desc(&mapMb); }
            uint32_t it = 0;
            RcWalk jw(w, nlanes);
            for (int64_t nj = 0; nj < myjobs; nj++, jw.advance(nwalk)) {
                int m0, n0, dir, chain;
                decode(jw.ln, m0, n0, dir, chain);
                const int src = (int)((jw.step - 1) & 1);
                RC2_TRACE(0, nj);
                const CUtensorMap *mapX = dir ? (src ? &mapXb1 : &mapXb0) : (src ? &mapXf1 : &mapXf0);
                const CUtensorMap *mapM = dir ? &mapMb : &mapMf;
                // k-block kb: arm the stage and fetch the matrix tile (it does not depend on the chain) / fetch the W tile
                auto load_m = [&](int kb, uint32_t itk) {
                    const int s = itk % STAGES;
                    mbar_wait(&bars->empty[s], ((itk / STAGES) & 1) ^ 1);
                    if (PAIR) {
                        // the leader alone arrives, expecting both CTAs' bytes; the peer's loads complete_tx on the leader's barrier
                        if (rank == 0) mbar_arrive_expect_tx(&bars->full[s], 2 * STAGE_BYTES);
                        tma_load_2d_pair(tiles + s * STAGE_BYTES + TC_TILE_BYTES, mapM, mapa_u32(smem_u32(&bars->full[s]), 0), kb * BKE,
                                         n0 + (int)rank * MROWS);
                    } else {
                        mbar_arrive_expect_tx(&bars->full[s], STAGE_BYTES);
                        tma_load_2d(tiles + s * STAGE_BYTES + TC_TILE_BYTES, mapM, &bars->full[s], kb * BKE, n0);
                    }
                };
                auto load_x = [&](int kb, uint32_t itk) {
                    const int s = itk % STAGES;
                    if (PAIR) tma_load_2d_pair(tiles + s * STAGE_BYTES, mapX, mapa_u32(smem_u32(&bars->full[s]), 0), kb * BKE, m0);
                    else tma_load_2d(tiles + s * STAGE_BYTES, mapX, &bars->full[s], kb * BKE, m0);
                };
                // the matrix tiles of the first STAGES k-blocks are requested BEFORE the dependency wait
                const int npre = nkb < STAGES ? nkb : STAGES;
                for (int kb = 0; kb < npre; kb++) load_m(kb, it + kb);
                // inputs: W_{step-1} of this CTA's chain = all its column-tile jobs of the previous step
                wait_chain(done + chain, per_step * (uint32_t)(jw.step - 1));
                RC2_TRACE(1, nj);
                asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy acquire -> TMA reads
                for (int kb = 0; kb < npre; kb++) load_x(kb, it + kb);
                for (int kb = npre; kb < nkb; kb++) { load_m(kb, it + kb); load_x(kb, it + kb); }
                it += nkb;
            }
        }
    } else if (warp == 9) {
        if (rank == 0 && elect_one()) {
            const uint32_t idesc = BF16 ? make_idesc_bf16(PAIR ? 256 : TC_BM, BN) : make_idesc_tf32(PAIR ? 256 : TC_BM, BN);
            uint32_t it = 0;
            for (int64_t nj = 0; nj < myjobs; nj++) {
                const int a = (int)(nj % NACC);
                mbar_wait(&bars->tmem_empty[a], (uint32_t)((nj / NACC) & 1) ^ 1);   // the epilogue(s) drained this accumulator
                RC2_TRACE(4, nj);
                tc_fence_after();
                for (int kb = 0; kb < nkb; kb++, it++) {
                    if (kb == 1) RC2_TRACE(5, nj);
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&bars->full[s], ph);
                    tc_fence_after();
                    const uint64_t da = make_kmajor_sw128_desc(tiles + s * STAGE_BYTES);
                    const uint64_t dbb = make_kmajor_sw128_desc(tiles + s * STAGE_BYTES + TC_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; k++) {      // 32 bytes of K per MMA: 8 tf32 or 16 bf16 elements
                        const uint64_t ak = desc_advance(da, k * 32), bk = desc_advance(dbb, k * 32);
                        const bool accum = (kb | k) != 0;
                        if (PAIR) { if (BF16) mma2_bf16_ss(tmem + a * BN, ak, bk, idesc, accum); else mma2_tf32_ss(tmem + a * BN, ak, bk, idesc, accum); }
                        else { if (BF16) mma_bf16_ss(tmem + a * BN, ak, bk, idesc, accum); else mma_tf32_ss(tmem + a * BN, ak, bk, idesc, accum); }
                    }
                    if (PAIR) tc_commit2(&bars->empty[s], 3); else tc_commit(&bars->empty[s]);
                }
                if (PAIR) tc_commit2(&bars->tmem_full[a], 3); else tc_commit(&bars->tmem_full[a]);
                RC2_TRACE(6, nj);
            }
        }
    } else {
        // epilogue: all eight warps work on EVERY job -- warp (wq, h) owns TMEM lanes 32 wq .. 32 wq + 31 (= sequences) and the
        // column half h of the tile, so a job's epilogue latency (which sits on the chain's critical path) is BN/64
        // 32-column chunks instead of BN/32.  Arithmetic is packed (FADD2/FMUL2/FFMA2): two states per instruction.
        const int h = warp >> 2, wq = warp & 3;
        RcWalk jw(w, nlanes);
        for (int64_t nj = 0; nj < myjobs; nj++, jw.advance(nwalk)) {
            const int a = (int)(nj % NACC);
            int m0, n0, dir, chain;
            decode(jw.ln, m0, n0, dir, chain);
            const TcDir d = dir ? db : df;
            const int src = (int)((jw.step - 1) & 1), dst = (int)(jw.step & 1);
            const int64_t b = m0 + wq * 32 + lane;
            const bool live = b < B;
            const int64_t t = dir ? (T - 1 - jw.step) : jw.step;
            const int64_t len_b = (RAGGED && live) ? (int64_t)em.lengths[b] : T;
            const bool pad = RAGGED && t >= len_b;            // a step of this row's padding: emission factor 1
            // backward pass of a padded row: w_t = 1 in the padding and w_{len-1} = b(o_{len-1}) EXACTLY, as if the
            // recursion started there (running the normalised constant vector through the padding instead would bias
            // log P: all its entries round to tf32 / bf16 the same way, +-2e-4 per step, measured)
            const bool restart = RAGGED && dir == 1 && t >= len_b - 1;
            float r = 0.f, x = 0.f;
            const float *brow = nullptr;
            if (live) {       // nothing here depends on the previous step
                if (GAUSS) {
                    x = load_real<float>(em.obs, em.obs_dtype, b * T + t);
                } else {
                    int sym = load_symbol(em.obs, em.obs_dtype, b * T + t);
                    sym = sym < 0 ? 0 : (sym >= em.M ? em.M - 1 : sym);
                    brow = em.p0 + (int64_t)sym * N + n0;
                }
            }
            if (warp == 0) RC2_TRACE(8, nj);
            // the previous step's row sums of this chain come from other CTAs: acquire them
            wait_chain(done + chain, per_step * (uint32_t)(jw.step - 1));
            if (live) {
                const float2 *ps = reinterpret_cast<const float2 *>(d.psum + (b * 2 + src) * NTq);
                float2 p2[8];
#pragma unroll
                for (int k = 0; k < 8; k++) p2[k] = (2 * k < NTq) ? __ldcg(ps + k) : make_float2(0.f, 0.f);   // N <= 1024: NTq <= 16
                const float c = (((p2[0].x + p2[0].y) + (p2[1].x + p2[1].y)) + ((p2[2].x + p2[2].y) + (p2[3].x + p2[3].y))) +
                                (((p2[4].x + p2[4].y) + (p2[5].x + p2[5].y)) + ((p2[6].x + p2[6].y) + (p2[7].x + p2[7].y)));
                r = 1.0f / c;
                // (ragged: the scale factors of padding steps are not part of log P -- forward: c_{t-1} with t-1 >= len;
                // backward: the sum of w_{t+1} with t+1 >= len)
                if (n0 == 0 && h == 0 && !(RAGGED && dir == 0 && t - 1 >= len_b) && !restart)
                    __stcg(d.logacc + b, __ldcg(d.logacc + b) + log((double)c));
            }
            if (warp == 0) RC2_TRACE(9, nj);
            mbar_wait(&bars->tmem_full[a], (uint32_t)((nj / NACC) & 1));
            if (warp == 0) RC2_TRACE(10, nj);
            tc_fence_after();
            const uint32_t tbase = tmem + ((uint32_t)(wq * 32) << 16) + a * BN + h * HALF;
            const int nh = n0 + h * HALF;                           // first column of this warp's half
            // W_t leaves through shared memory: thread = row writes its 16-byte pieces into a [128 rows][128 B] box with the
            // TMA 128-byte swizzle (piece p of row r at p ^ (r & 7): conflict-free), one thread stores the box with a tensor
            // store -- full 128-byte lines to L2 instead of 16-byte pieces of 32 different lines per warp instruction.
            unsigned char *obuf = stage_out + h * TC_TILE_BYTES;
            const int orow = wq * 32 + lane;
            unsigned char *orowp = obuf + orow * 128;
            const CUtensorMap *mapO = dir ? (dst ? &mapXb1 : &mapXb0) : (dst ? &mapXf1 : &mapXf0);
            constexpr int BOXC = BF16 ? 64 : 32;                   // columns per 128-byte box row
            const float2 rr = make_float2(r, r), xx = make_float2(x, x);
            float2 rs0 = make_float2(0.f, 0.f), rs1 = make_float2(0.f, 0.f);
            auto process = [&](const uint32_t (&acc)[32], int c0) {       // c0: column inside this warp's half
                float2 v[16];
                if (!live) {
#pragma unroll
                    for (int q = 0; q < 16; q++) v[q] = make_float2(0.f, 0.f);
                } else if (GAUSS) {
                    const float4 *nm4 = reinterpret_cast<const float4 *>(sp + nh + c0);
                    const float4 *k4 = reinterpret_cast<const float4 *>(sp + N + nh + c0);
                    const float4 *l4 = reinterpret_cast<const float4 *>(sp + 2 * N + nh + c0);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float4 nm = nm4[q], kk = k4[q], ll = l4[q];
                        const float2 d0 = __fadd2_rn(xx, make_float2(nm.x, nm.y)), d1 = __fadd2_rn(xx, make_float2(nm.z, nm.w));
                        const float2 a0 = __ffma2_rn(__fmul2_rn(d0, d0), make_float2(kk.x, kk.y), make_float2(ll.x, ll.y));
                        const float2 a1 = __ffma2_rn(__fmul2_rn(d1, d1), make_float2(kk.z, kk.w), make_float2(ll.z, ll.w));
                        float e0, e1, e2, e3;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(pad ? 0.f : a0.x));
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(pad ? 0.f : a0.y));
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(pad ? 0.f : a1.x));
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(pad ? 0.f : a1.y));
                        float2 ar0 = __fmul2_rn(make_float2(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1])), rr);
                        float2 ar1 = __fmul2_rn(make_float2(__uint_as_float(acc[4 * q + 2]), __uint_as_float(acc[4 * q + 3])), rr);
                        if (restart) { ar0 = make_float2(1.f, 1.f); ar1 = ar0; }
                        v[2 * q] = __fmul2_rn(ar0, make_float2(e0, e1));
                        v[2 * q + 1] = __fmul2_rn(ar1, make_float2(e2, e3));
                    }
                } else {
                    const float4 *b4 = reinterpret_cast<const float4 *>(brow + h * HALF + c0);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        float4 e = __ldg(b4 + q);
                        if (pad) e = make_float4(1.f, 1.f, 1.f, 1.f);
                        float2 ar0 = __fmul2_rn(make_float2(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1])), rr);
                        float2 ar1 = __fmul2_rn(make_float2(__uint_as_float(acc[4 * q + 2]), __uint_as_float(acc[4 * q + 3])), rr);
                        if (restart) { ar0 = make_float2(1.f, 1.f); ar1 = ar0; }
                        v[2 * q] = __fmul2_rn(ar0, make_float2(e.x, e.y));
                        v[2 * q + 1] = __fmul2_rn(ar1, make_float2(e.z, e.w));
                    }
                }
                if (c0 == 64) {       // (BN = 256) the first 64 columns of this half are one row-sum slot
                    if (live) d.psum[(b * 2 + dst) * NTq + (nh >> 6)] = (rs0.x + rs0.y) + (rs1.x + rs1.y);
                    rs0 = make_float2(0.f, 0.f);
                    rs1 = make_float2(0.f, 0.f);
                }
#pragma unroll
                for (int q = 0; q < 8; q++) { rs0 = __fadd2_rn(rs0, v[2 * q]); rs1 = __fadd2_rn(rs1, v[2 * q + 1]); }
                if ((c0 % BOXC) == 0) {
                    // the box of the previous group must have been read out by its tensor store
                    if (threadIdx.x == h * 128) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
                }
                if (BF16) {
                    const int p0 = (c0 & 32) ? 4 : 0;               // second 32-column half of the 64-column box
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        *reinterpret_cast<uint4 *>(orowp + (((p0 + q) ^ (orow & 7)) << 4)) =
                            make_uint4(pack_bf16(v[4 * q].x, v[4 * q].y), pack_bf16(v[4 * q + 1].x, v[4 * q + 1].y),
                                       pack_bf16(v[4 * q + 2].x, v[4 * q + 2].y), pack_bf16(v[4 * q + 3].x, v[4 * q + 3].y));
                } else {
#pragma unroll
                    for (int q = 0; q < 8; q++)
                        *reinterpret_cast<float4 *>(orowp + ((q ^ (orow & 7)) << 4)) =
                            make_float4(round_tf32(v[2 * q].x), round_tf32(v[2 * q].y), round_tf32(v[2 * q + 1].x), round_tf32(v[2 * q + 1].y));
                }
                if (((c0 + 32) % BOXC) == 0) {
                    fence_proxy_async();                             // generic-proxy smem writes -> async-proxy (TMA) reads
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
                    if (threadIdx.x == h * 128) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(mapO), "r"(smem_u32(obuf)), "r"(nh + c0 + 32 - BOXC), "r"(m0) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            };
            uint32_t accA[32], accB[32];
            tmem_ld_32x32(tbase, accA);
#pragma unroll 1
            for (int c0 = 0; c0 < HALF; c0 += 64) {
                tmem_ld_wait();
                if (c0 == 0 && warp == 0) RC2_TRACE(12, nj);
                tmem_ld_32x32(tbase + c0 + 32, accB);          // in flight under the arithmetic on accA
                process(accA, c0);
                if (c0 == 0 && warp == 0) RC2_TRACE(13, nj);
                tmem_ld_wait();
                if (c0 + 64 < HALF) tmem_ld_32x32(tbase + c0 + 64, accA);
                process(accB, c0 + 32);
            }
            if (warp == 0) RC2_TRACE(14, nj);
            if (live) d.psum[(b * 2 + dst) * NTq + ((nh + HALF - 64) >> 6)] = (rs0.x + rs0.y) + (rs1.x + rs1.y);
            // accumulator a may be overwritten: one arrival per epilogue warp, on the LEADER's barrier
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->tmem_empty[a]), 0));
                else mbar_arrive(&bars->tmem_empty[a]);
            }
            // publish this half of the job: the barrier orders the 128 threads' row-sum / log-accumulator stores before
            // the elected thread, whose gpu-scope release is cumulative over them (the pattern of a grid barrier); its
            // own tensor stores are complete (wait_group) first
            asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
            if (warp == 0) RC2_TRACE(15, nj);
            if (threadIdx.x == h * 128) {
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // complete = visible to the generic proxy (implicit proxy fence)
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(done + chain) : "memory");
            }
            if (warp == 0) RC2_TRACE(11, nj);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();      // the peer may still signal our barriers / use the paired TMEM
    if (warp == 9) { if (PAIR) tmem_dealloc2(tmem, 512); else tmem_dealloc(tmem, 512); }
#undef RC2_TRACE
}

template <bool GAUSS, bool BF16, int BN, bool RAGGED>
__global__ void __launch_bounds__(320, 1)
tc_recursion2_kernel(const __grid_constant__ CUtensorMap mapXf0, const __grid_constant__ CUtensorMap mapXf1,
                     const __grid_constant__ CUtensorMap mapXb0, const __grid_constant__ CUtensorMap mapXb1,
                     const __grid_constant__ CUtensorMap mapMf, const __grid_constant__ CUtensorMap mapMb, TcDir df, TcDir db,
                     TcEmis em, int64_t B, int64_t T, int N, int mtiles, int ndir, uint32_t *done, long long *trace) {
    tc_recursion2_body<GAUSS, BF16, false, BN, RAGGED>(mapXf0, mapXf1, mapXb0, mapXb1, mapMf, mapMb, df, db, em, B, T, N, mtiles, ndir, done, trace);
}
template <bool GAUSS, bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
tc_recursion2_pair_kernel(const __grid_constant__ CUtensorMap mapXf0, const __grid_constant__ CUtensorMap mapXf1,
                          const __grid_constant__ CUtensorMap mapXb0, const __grid_constant__ CUtensorMap mapXb1,
                          const __grid_constant__ CUtensorMap mapMf, const __grid_constant__ CUtensorMap mapMb, TcDir df, TcDir db,
                          TcEmis em, int64_t B, int64_t T, int N, int mtiles, int ndir, uint32_t *done, long long *trace) {
    tc_recursion2_body<GAUSS, BF16, true, 256, false>(mapXf0, mapXf1, mapXb0, mapXb1, mapMf, mapMb, df, db, em, B, T, N, mtiles, ndir, done, trace);
}

// step 0: W_0 = start . b(o_first): start = pi (forward, t = 0) or ones (backward, t = T-1)
template <bool GAUSS, bool BF16 = false>
__global__ void tc_init_kernel(TcDir df, TcDir db, TcEmis em, const float *__restrict__ pi, int64_t B, int64_t T, int N, int psg) {
    // psg: columns per partial-row-sum slot (128 for the per-step kernels, 64 for the whole-recursion kernel)
    const int dir = blockIdx.y;
    const TcDir d = dir ? db : df;
    const int64_t b = blockIdx.x;
    const int NT = N / TC_BN;
    const int64_t t = dir ? T - 1 : 0;
    const bool pad = em.lengths && t >= em.lengths[b];      // (backward start of a padded row: emission factor 1)
    float x = 0.f;
    const float *brow = nullptr;
    if (GAUSS) {
        x = load_real<float>(em.obs, em.obs_dtype, b * T + t);
    } else {
        int sym = load_symbol(em.obs, em.obs_dtype, b * T + t);
        sym = sym < 0 ? 0 : (sym >= em.M ? em.M - 1 : sym);
        brow = em.p0 + (int64_t)sym * N;
    }
    __shared__ float red[32];
    for (int nt = 0; nt < NT; nt++) {
        int j = nt * TC_BN + threadIdx.x;  // blockDim.x == 128
        float e;
        if (GAUSS) {
            float sd = em.p1[j], dd = x - em.p0[j];
            e = exp2f(fmaf(dd * dd, -0.5f * 1.4426950408889634f / (sd * sd), -log2f(sd * 2.5066282746310002f)));
        } else {
            e = brow[j];
        }
        if (pad) e = 1.f;
        float v = (dir ? 1.f : pi[j]) * e;
        if (BF16) reinterpret_cast<__nv_bfloat16 *>(d.W)[(b * 2 + 0) * N + j] = __float2bfloat16_rn(v);
        else d.W[(b * 2 + 0) * N + j] = round_tf32(v);
        float s = warp_sum(v);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            if (psg == 64) {
                d.psum[(b * 2 + 0) * (2 * NT) + 2 * nt] = red[0] + red[1];
                d.psum[(b * 2 + 0) * (2 * NT) + 2 * nt + 1] = red[2] + red[3];
            } else {
                d.psum[(b * 2 + 0) * NT + nt] = red[0] + red[1] + red[2] + red[3];
            }
        }
    }
    if (threadIdx.x == 0) d.logacc[b] = 0.0;
}

// closing: forward log P = logacc + log c_{T-1}; backward log P = logacc + log(pi . w_0)
template <bool BF16 = false>
__global__ void tc_final_kernel(TcDir df, TcDir db, const float *__restrict__ pi, int64_t B, int64_t T, int N,
                                const int32_t *__restrict__ lengths, double *__restrict__ loglik,
                                double *__restrict__ loglik_bwd) {
    const int dir = blockIdx.y;
    const TcDir d = dir ? db : df;
    const int64_t b = blockIdx.x;
    const int slot = (int)((T - 1) & 1);
    const float *w = d.W + (b * 2 + slot) * N;
    const __nv_bfloat16 *wh = reinterpret_cast<const __nv_bfloat16 *>(d.W) + (b * 2 + slot) * N;
    float s = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const float wj = BF16 ? __bfloat162float(wh[j]) : w[j];
        s += dir ? wj * pi[j] : wj;
    }
    __shared__ float red[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) tot += red[k];
        // (forward, ragged: the last step of a padded row is padding -- its c_{T-1} is not part of log P)
        const bool skip_last = dir == 0 && lengths && lengths[b] < T;
        double ll = d.logacc[b] + (skip_last ? 0.0 : log((double)tot));
        if (dir) { if (loglik_bwd) loglik_bwd[b] = ll; } else loglik[b] = ll;
    }
}

static size_t tc_workspace_floats(int64_t B, int N) {
    int NT = N / TC_BN;
    int64_t Bp = (B + TC_BM - 1) / TC_BM * TC_BM;
    // per direction: W [Bp][2][N], psum [Bp][2][NT], logacc [Bp] doubles (2 floats each)
    // + per-chain completion counters of the whole-recursion kernel (2 directions x row tiles)
    // (row sums per 64 columns in the whole-recursion kernel: 2 NT slots)
    return (size_t)2 * ((size_t)Bp * 2 * N + (size_t)Bp * 2 * 2 * NT + (size_t)Bp * 2) + (size_t)2 * N * N + 2 * (size_t)(Bp / TC_BM) + 64;
}

static long long *g_rc_trace = nullptr;

template <bool GAUSS>
static int run_tc(int64_t B, int64_t T, int N, const float *pi, const float *A, const float *AT, TcEmis em, void *ws,
                  size_t ws_bytes, double *loglik, double *loglik_bwd, int precision, cudaStream_t st, bool no_pairs = false) {
    KHMM_REQUIRE(precision == KHMM_TC_TF32 || precision == KHMM_TC_BF16, "fwdbwd_loglik_tc: unknown operand precision %d", precision);
    const bool bf16 = precision == KHMM_TC_BF16;
    KHMM_REQUIRE(B >= 0 && T >= 1, "fwdbwd_loglik_tc: bad shape");
    if (N < TC_BN || N % TC_BN != 0 || N > 1024) {
        set_error("fwdbwd_loglik_tc: N=%d must be a multiple of %d in [%d, 1024]", N, TC_BN, TC_BN);
        return KHMM_ERR_UNSUPPORTED;
    }
    if (B == 0) return KHMM_OK;
    KHMM_REQUIRE(pi && A && AT && ws && loglik, "fwdbwd_loglik_tc: NULL pointer argument");
    KHMM_REQUIRE(ws_bytes >= tc_workspace_floats(B, N) * sizeof(float), "fwdbwd_loglik_tc: workspace too small");
    KHMM_REQUIRE(((uintptr_t)ws & 255) == 0, "fwdbwd_loglik_tc: workspace must be 256-byte aligned");
    const int NT = N / TC_BN;
    const int64_t Bp = (B + TC_BM - 1) / TC_BM * TC_BM;
    float *p = (float *)ws;
    TcDir d[2];
    for (int k = 0; k < 2; k++) {
        d[k].W = p;      p += (size_t)Bp * 2 * N;
        d[k].psum = p;   p += (size_t)Bp * 2 * 2 * NT;
        d[k].logacc = (double *)p; p += (size_t)Bp * 2;
    }
    float *Ar = p;  p += (size_t)N * N;     // tf32-rounded copies of A and AT
    float *ATr = p; p += (size_t)N * N;
    uint32_t *done = (uint32_t *)p; p += 2 * (size_t)(Bp / TC_BM);
    const int ndir = loglik_bwd ? 2 : 1;
    int coop = 0;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    }
    // KHMM_RC_MODE (development switch): 1 = one launch per step (the path without cooperative launch), 2 = whole-recursion
    // kernel on single CTAs, 3 = on CTA pairs where the shape allows.  Measured at c5 (B = 4096, T = 1024, N = 1024): bf16 17.5 ms
    // single / 17.9 ms pairs; tf32 ~29 ms single / 25.8 ms pairs
    int rc_mode = KHMM_RC2 ? ((KHMM_RC2_PAIR && !bf16) ? 3 : 2) : 1;
    if (dev_switch(SW_RC_MODE) >= 0) rc_mode = dev_switch(SW_RC_MODE);
    if (no_pairs && rc_mode == 3) rc_mode = 2;
    const bool use_rc2 = coop && rc_mode >= 2;         // any N % 128 == 0, any batch: one launch for all T-1 steps
    if (em.lengths && !use_rc2) {
        set_error("fwdbwd_loglik_tc: ragged batches need the whole-recursion kernel (cooperative launch)");
        return KHMM_ERR_UNSUPPORTED;
    }
    if (bf16) {
        if (!use_rc2) {
            set_error("fwdbwd_loglik_tc: bf16 operands need the whole-recursion kernel (cooperative launch)");
            return KHMM_ERR_UNSUPPORTED;
        }
        tc_to_bf16_kernel<<<64, 256, 0, st>>>(A, (__nv_bfloat16 *)Ar, N * N);
        tc_to_bf16_kernel<<<64, 256, 0, st>>>(AT, (__nv_bfloat16 *)ATr, N * N);
        tc_init_kernel<GAUSS, true><<<dim3((unsigned)B, ndir), 128, 0, st>>>(d[0], d[1], em, pi, B, T, N, use_rc2 ? 64 : 128);
    } else {
        tc_round_matrix_kernel<<<64, 256, 0, st>>>(A, Ar, N * N);
        tc_round_matrix_kernel<<<64, 256, 0, st>>>(AT, ATr, N * N);
        tc_init_kernel<GAUSS, false><<<dim3((unsigned)B, ndir), 128, 0, st>>>(d[0], d[1], em, pi, B, T, N, use_rc2 ? 64 : 128);
    }
    KHMM_LAUNCH_CHECK("tc_init_kernel");
    // tensor maps: the W buffer [Bp][2][N] is addressed as a 2-D matrix whose row b is (b, slot):
    // base = W + slot*N, row stride = 2*N floats -> one map per (direction, slot)
    CUtensorMap mapX[2][2], mapM[2];
    for (int k = 0; k < 2; k++)
        for (int slot = 0; slot < 2; slot++) {
            // rows = Bp, cols = N, but stride 2N: encode as matrix with cols = 2N and only the first N used
            static PFN_encodeTiled encode = nullptr;
            if (!encode) {
                void *fn = nullptr;
                cudaDriverEntryPointQueryResult q;
                if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
                    set_error("cuTensorMapEncodeTiled entry point not found");
                    return KHMM_ERR_CUDA;
                }
                encode = (PFN_encodeTiled)fn;
            }
            const size_t esz = bf16 ? 2 : 4;
            cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)Bp};
            cuuint64_t strides[1] = {(cuuint64_t)2 * N * esz};
            cuuint32_t box[2] = {(cuuint32_t)(bf16 ? 64 : TC_BK), TC_BM};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = encode(&mapX[k][slot], bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                                (void *)((char *)d[k].W + (size_t)slot * N * esz),
                                dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
                return KHMM_ERR_CUDA;
            }
        }
    // 256-wide column tiles when N allows: 25 % less L2 operand traffic per MAC and half the CTAs
    // (small batches -- e.g. a strong-scaling shard -- keep 128-wide tiles so that more SMs get a CTA)
    int BN = (N % 256 == 0 && (Bp / TC_BM) * (N / 256) * ndir >= sm_count()) ? 256 : 128;
    if (dev_switch(SW_RC_BN) == 128) BN = 128;      // development switch
    // forward: out[j] = sum_i W[i] A[i][j]  -> B-operand rows j, K index i: AT[j][i]
    // backward: out[i] = sum_j A[i][j] w[j] -> B-operand rows i, K index j: A[i][j]
    if (bf16 ? (make_tmap_bf16_rowmajor(&mapM[0], ATr, N, N, BN) || make_tmap_bf16_rowmajor(&mapM[1], Ar, N, N, BN))
             : (make_tmap_f32_rowmajor(&mapM[0], ATr, N, N, BN) || make_tmap_f32_rowmajor(&mapM[1], Ar, N, N, BN))) {
        set_error("tensor map creation for the transition matrix failed");
        return KHMM_ERR_CUDA;
    }
    if (use_rc2) {
        const int mtiles = (int)(Bp / TC_BM);
        const bool ragged = em.lengths != nullptr;
        const bool rpair = rc_mode == 3 && BN == 256 && mtiles % 2 == 0 && !ragged;      // (the pair kernel has no ragged variant)
        CUtensorMap mapH[2];
        if (rpair && (bf16 ? (make_tmap_bf16_rowmajor(&mapH[0], ATr, N, N, 128) || make_tmap_bf16_rowmajor(&mapH[1], Ar, N, N, 128))
                           : (make_tmap_f32_rowmajor(&mapH[0], ATr, N, N, 128) || make_tmap_f32_rowmajor(&mapH[1], Ar, N, N, 128)))) {
            set_error("tensor map creation for the transition matrix failed");
            return KHMM_ERR_CUDA;
        }
        const size_t ring = rpair ? (size_t)TcRBars<true, 256>::STAGES * (TC_TILE_BYTES + 128 * 128)
                                  : (BN == 256 ? (size_t)TcRBars<false, 256>::STAGES * (TC_TILE_BYTES + 256 * 128)
                                               : (size_t)TcRBars<false, 128>::STAGES * (TC_TILE_BYTES + 128 * 128));
        const size_t smem = ring + 2 * (size_t)TC_TILE_BYTES + sizeof(TcRBars<true, 256>) + 16 + 3 * (size_t)N * sizeof(float) + 1024;
        void *kern = rpair ? (bf16 ? (void *)tc_recursion2_pair_kernel<GAUSS, true> : (void *)tc_recursion2_pair_kernel<GAUSS, false>)
                   : ragged ? (BN == 256 ? (bf16 ? (void *)tc_recursion2_kernel<GAUSS, true, 256, true> : (void *)tc_recursion2_kernel<GAUSS, false, 256, true>)
                                         : (bf16 ? (void *)tc_recursion2_kernel<GAUSS, true, 128, true> : (void *)tc_recursion2_kernel<GAUSS, false, 128, true>))
                   : BN == 256 ? (bf16 ? (void *)tc_recursion2_kernel<GAUSS, true, 256, false> : (void *)tc_recursion2_kernel<GAUSS, false, 256, false>)
                               : (bf16 ? (void *)tc_recursion2_kernel<GAUSS, true, 128, false> : (void *)tc_recursion2_kernel<GAUSS, false, 128, false>);
        KHMM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int nlanes = (rpair ? mtiles / 2 : mtiles) * (N / BN) * ndir;
        int walkers = rpair ? sm_count() / 2 : sm_count();
        if (rpair) {
            // every CTA pair must be resident at once: ask how many clusters of two the device can hold
            cudaLaunchConfig_t qc = {};
            qc.gridDim = dim3(2 * walkers);
            qc.blockDim = dim3(320);
            qc.dynamicSmemBytes = smem;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &qc) != cudaSuccess || nclusters < 1) {
                cudaGetLastError();
                nclusters = walkers;
            }
            if (nclusters < walkers) walkers = nclusters;
        }
        if (nlanes < walkers) walkers = nlanes;
        const int grid = rpair ? 2 * walkers : walkers;
        KHMM_CUDA_CHECK(cudaMemsetAsync(done, 0, 2 * (size_t)mtiles * sizeof(uint32_t), st));
        int Ni = N, mt = mtiles, nd = ndir;
        int64_t Bv = B, Tv = T;
        void *args[] = {&mapX[0][0], &mapX[0][1], &mapX[1][0], &mapX[1][1], rpair ? &mapH[0] : &mapM[0], rpair ? &mapH[1] : &mapM[1],
                        &d[0], &d[1], &em, &Bv, &Tv, &Ni, &mt, &nd, &done, &g_rc_trace};
        cudaError_t le = cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(320), args, smem, st);
        if (le != cudaSuccess && rpair) {
            // cooperative launch of the cluster kernel refused: take the single-CTA whole-recursion kernel (also a
            // cooperative launch: a plain launch of kernels that spin on each other's progress could hang when the
            // CTAs are not all resident, e.g. under MPS or with other streams active)
            cudaGetLastError();
            return run_tc<GAUSS>(B, T, N, pi, A, AT, em, ws, ws_bytes, loglik, loglik_bwd, precision, st, true);
        }
        KHMM_CUDA_CHECK(le);
    } else {
        // no cooperative launch (or KHMM_RC_MODE=1): one launch per time step, tf32 operands
        size_t smem = TC_STAGES * (size_t)(TC_TILE_BYTES + BN * TC_BK * 4) + sizeof(TcBars) + 3 * BN * sizeof(float) + 1024;
        auto kern = BN == 256 ? tc_step_kernel<GAUSS, 256> : tc_step_kernel<GAUSS, 128>;
        KHMM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)(Bp / TC_BM), N / BN, ndir);
        for (int step = 1; step < T; step++) {
            int src = (step - 1) & 1;
            kern<<<grid, 192, smem, st>>>(mapX[0][src], mapX[1][src], mapM[0], mapM[1], d[0], d[1], em, B, T, N, step);
        }
        KHMM_LAUNCH_CHECK("tc_step_kernel");
    }
    if (bf16) tc_final_kernel<true><<<dim3((unsigned)B, ndir), 256, 0, st>>>(d[0], d[1], pi, B, T, N, em.lengths, loglik, loglik_bwd);
    else tc_final_kernel<false><<<dim3((unsigned)B, ndir), 256, 0, st>>>(d[0], d[1], pi, B, T, N, em.lengths, loglik, loglik_bwd);
    KHMM_LAUNCH_CHECK("tc_final_kernel");
    return KHMM_OK;
}

}  // namespace khmm

using namespace khmm;

extern "C" {

int khmm_debug_rc_trace(int64_t *dev_buf) {
    g_rc_trace = (long long *)dev_buf;
    return KHMM_OK;
}

size_t khmm_fwdbwd_loglik_tc_workspace_bytes(int64_t B, int N) {
    if (N < TC_BN || N % TC_BN != 0) return 0;
    return tc_workspace_floats(B, N) * sizeof(float);
}

int khmm_fwdbwd_loglik_tc_gauss_ragged_f32(int64_t B, int64_t T, int N, const float *pi, const float *A, const float *AT,
                                           const float *mean, const float *sd, const void *obs, int obs_dtype,
                                           const int32_t *lengths, void *workspace, size_t workspace_bytes, double *loglik,
                                           double *loglik_bwd, int operand_precision, khmm_stream_t stream) {
    KHMM_REQUIRE(B == 0 || (mean && sd && obs), "fwdbwd_loglik_tc_gauss: NULL pointer argument");
    KHMM_REQUIRE(obs_dtype == KHMM_F32 || obs_dtype == KHMM_F64, "gauss: obs dtype code %d is not f32/f64", obs_dtype);
    TcEmis em{1, mean, sd, obs, obs_dtype, 0, lengths};
    return run_tc<true>(B, T, N, pi, A, AT, em, workspace, workspace_bytes, loglik, loglik_bwd, operand_precision,
                        (cudaStream_t)stream);
}

int khmm_fwdbwd_loglik_tc_gauss_f32(int64_t B, int64_t T, int N, const float *pi, const float *A, const float *AT,
                                    const float *mean, const float *sd, const void *obs, int obs_dtype, void *workspace,
                                    size_t workspace_bytes, double *loglik, double *loglik_bwd, int operand_precision,
                                    khmm_stream_t stream) {
    return khmm_fwdbwd_loglik_tc_gauss_ragged_f32(B, T, N, pi, A, AT, mean, sd, obs, obs_dtype, nullptr, workspace,
                                                  workspace_bytes, loglik, loglik_bwd, operand_precision, stream);
}

int khmm_fwdbwd_loglik_tc_discrete_ragged_f32(int64_t B, int64_t T, int N, int M, const float *pi, const float *A,
                                              const float *AT, const float *Bsm, const void *obs, int obs_dtype,
                                              const int32_t *lengths, void *workspace, size_t workspace_bytes,
                                              double *loglik, double *loglik_bwd, int operand_precision,
                                              khmm_stream_t stream) {
    KHMM_REQUIRE(B == 0 || (Bsm && obs && M >= 1), "fwdbwd_loglik_tc_discrete: NULL pointer argument");
    KHMM_REQUIRE(obs_dtype >= KHMM_U8 && obs_dtype <= KHMM_I64, "discrete: obs dtype code %d is not an integer type",
                 obs_dtype);
    TcEmis em{0, Bsm, nullptr, obs, obs_dtype, M, lengths};
    return run_tc<false>(B, T, N, pi, A, AT, em, workspace, workspace_bytes, loglik, loglik_bwd, operand_precision,
                         (cudaStream_t)stream);
}

int khmm_fwdbwd_loglik_tc_discrete_f32(int64_t B, int64_t T, int N, int M, const float *pi, const float *A,
                                       const float *AT, const float *Bsm, const void *obs, int obs_dtype,
                                       void *workspace, size_t workspace_bytes, double *loglik, double *loglik_bwd,
                                       int operand_precision, khmm_stream_t stream) {
    return khmm_fwdbwd_loglik_tc_discrete_ragged_f32(B, T, N, M, pi, A, AT, Bsm, obs, obs_dtype, nullptr, workspace,
                                                     workspace_bytes, loglik, loglik_bwd, operand_precision, stream);
}

}  // extern "C"
