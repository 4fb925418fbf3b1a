// Tensor-core Baum-Welch E-step for N = 128 states (config 4: DiscreteHMM 128 x 1024, T = 512).
//
// Two persistent tcgen05 kernels per batch chunk replace forward (abstract_hmm.py:250-285), backward
// (:327-351), gamma (:291-297), zi (:299-325) and the count sums of DiscreteHMM.do_pass
// (discrete_hmm.py:117-159):
//
//   bw_forward_tc_kernel   thread = sequence.  CTA = one tile of 128 sequences, whole time loop inside the kernel.  The row
//       W_t = c_t alpha_t (unnormalised: its row sum IS the scale factor c_t) stays in TENSOR MEMORY as
//       the A operand of the next step's MMA (tcgen05.mma with A in TMEM, A^T as a K-major B operand in
//       shared memory, loaded once by TMA); the epilogue multiplies the accumulator by the gathered emission row and by
//       1/c_{t-1}, writes W_t back to TMEM and to the HBM stash [tile][T][128][128] (fp32, TMA tensor stores), and keeps
//       c_t / log P.  The MMA of a step is issued as two column halves so that the epilogue of the first overlaps the second.
//   bw_backward_tc3_kernel  thread = STATE (the default backward kernel, round 2).  A tile is four quarter-tiles of 32
//       sequences that run as independent pipelines (sixteen compute warps):
//         MMA1  D1^T = A . V2_{t+1}^T        (A resident in TMEM, V2 K-major in shared memory, N = 32)
//         MMA2  Xi^T += V2_{t+1}^T . W_t     (V2^T in TMEM, W_t straight from the stash by TMA; Xi = zi summed over t, in TMEM)
//         threads: beta_t = D1 c_t / c_{t+1}; gamma_ref = W_t . beta_t -> red.global.add into row o_t of the fp32 count
//                  table, one coalesced 128-byte row per warp instruction straight from registers; V2_t = b(o_t) . beta_t /
//                  c_{t-1} -> shared memory and TMEM; emission values read from the table one step ahead (coalesced).
//   bw_backward_tc_kernel  thread = sequence (round 1): the independent cross-check of the kernel above (development switch).
//   bw_fold_kernel adds the fp32 table replicas into the float64 count buffer (include/khmm.h layout).
//
// HBM traffic per (sequence, step): 512 B stash write + 512 B stash read + 6 B observations/scales
// (the generic path moves 2.5 KB).  Operand precision of the recursion's contraction: tf32, or bf16 hi/lo pairs with
// three MMAs (KHMM_TC_BF16X3, fp32-class parity; the API default); fp32 accumulation; tests state the bounds.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "tcgen05.cuh"

namespace khmm {

using namespace tc;

constexpr int BW_N = 128;
constexpr int BW_SUB = 128 * 32 * 4;   // one [128 rows][32 fp32] sub-tile (16 KB)
constexpr int BW_TILE = 4 * BW_SUB;    // 128 x 128 fp32 operand tile (64 KB)
constexpr int BW_REPL = 8;             // fp32 count-table replicas (shorter fp32 accumulation chains)
// TMEM columns of the backward kernel
constexpr uint32_t COL_D1 = 0, COL_V = 128, COL_XI = 256, COL_STG = 384;

__device__ __forceinline__ float bw_round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float *p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
// 256-bit read-only load (sm_100): one full 32-byte sector per lane, half the L1 wavefronts of two LDG.128
__device__ __forceinline__ void ldg8(const float *p, float *o) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7])
                 : "l"(p));
}

static __global__ void bw_round_matrix_kernel(const float *__restrict__ in, float *__restrict__ out, int n) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) out[k] = bw_round_tf32(in[k]);
}

// bf16 hi/lo split of a matrix for the split-operand mode: hi = bf16(x), lo = bf16(x - hi)
static __global__ void bw_split_matrix_kernel(const float *__restrict__ in, uint16_t *__restrict__ hi, uint16_t *__restrict__ lo,
                                              int n) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        uint32_t h, l;
        split_bf16x2(in[k], 0.f, h, l);
        hi[k] = (uint16_t)(h & 0xffffu);
        lo[k] = (uint16_t)(l & 0xffffu);
    }
}

struct BwObs {
    const void *obs;
    int dtype, M;
    int64_t T;
    const int32_t *lengths;    // NULL: every sequence has T steps; else 1 <= lengths[b] <= T (ragged batch, rows padded to T)
    __device__ __forceinline__ int sym(int64_t b, int64_t t) const {
        int s = load_symbol(obs, dtype, b * T + t);
        return s < 0 ? 0 : (s >= M ? M - 1 : s);
    }
    __device__ __forceinline__ int len(int64_t b) const { return lengths ? lengths[b] : (int)T; }
};
// Ragged batches cost nothing extra: the recursions simply run on through a sequence's padding (its symbols index valid
// table rows, every value stays finite) and three per-row switches remove the padding from the statistics -- the forward
// kernel stashes ZERO rows for t >= len (so gamma_t = W_t . beta_t and the W_t^T V2_{t+1} products of those steps vanish)
// and leaves their c_t out of log P; the backward kernel sets beta_t = 1 for every t >= len - 1 (the reference's start value,
// abstract_hmm.py:337) and V2_t = 0 for t >= len (the pair (len-1, len) must not count); the reduction warps take a row's
// "last step" at t = len - 1 and skip its padding.

// ------------------------------------------------------------------------------- forward
// The first versions of this kernel did the emission gather and the stash store thread-per-row (each LDG/STG
// instruction touches 32 different lines: the L1 wavefront path made a step cost 12,000 cycles) and then with
// one bulk copy per thread (UBLKCP takes uniform operands, so a warp serialises 32 of them: 40 % of the time).
// Now dedicated warps move rows with fully coalesced accesses -- lane = 16-byte piece of ONE row -- and the
// epilogue warps only compute.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// arrive on the mbarrier once all cp.async issued so far by this thread have landed (counts as one of the
// barrier's expected arrivals)
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// round to nearest tf32 for finite positive values: add half an ulp of the 10-bit mantissa, clear 13 bits
__device__ __forceinline__ uint32_t tf32_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

constexpr int FW_EBUF = BW_TILE;   // one emission / stash staging buffer: 4 sub-tiles [128 rows][128 B], SWIZZLE_128B

struct __align__(8) FwdBars {
    uint64_t mat_full, a_ready, d_full[2], e_full[2], w_ready[2], buf_free[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *smem_src, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Shared memory: A^T (64 KB, K-major SW128) + two row buffers in the TMA box layout (4 sub-tiles of
// [128 rows][32 floats], 16-byte pieces XOR (row & 7): thread-per-row LDS/STS.128 is conflict-free).  Buffer t&1:
// the gather warps copy the emission rows b(o_t) of the tile into it one step ahead (cp.async, lane = 16-byte
// piece, 4 rows per instruction), the epilogue overwrites them IN PLACE with W_t, and ONE thread sends the
// whole 64 KB tile to the tile-major stash with four TMA tensor stores (plus one 512-byte bulk store for c_t).
// Measured history of this kernel (clock64 timeline, tools/estep_forward_timeline.py): thread-per-row LDG/STG 12,000
// cycles per step (32 L1 wavefronts per instruction); per-thread bulk copies 10,000 (UBLKCP serialises 32 per
// warp); a storer warp doing LDS+STG per row 5,200 (the storer alone took 4,500).
// Warps: 0-7 epilogue (thread = sequence = TMEM lane; warps 0-3 own columns 0-63, warps 4-7 columns 64-127 of
// the same rows, the two halves of a row sum meet in shared memory), 8 MMA issuer + TMEM allocator, 9-10 gather,
// 11 storer.
// SPLIT: the recursion's MMA takes bf16 hi/lo operand pairs (kind::f16) instead of tf32 -- W_hi, W_lo in TMEM (two
// values per column), A^T_hi, A^T_lo in the same 64 KB of shared memory -- and adds the three products W_hi A_hi +
// W_hi A_lo + W_lo A_hi (2^-17 operand precision against tf32's 2^-11, 1.5 x the tensor time): the rounding of the
// transition matrix, which tf32 leaves as a SYSTEMATIC 1e-5 ... 1e-4 error in the counts, drops below 2e-6.
template <int PREC>
__global__ void __launch_bounds__(384, 1)
bw_forward_tc_kernel(const __grid_constant__ CUtensorMap mapAT, const __grid_constant__ CUtensorMap mapATlo,
                     const __grid_constant__ CUtensorMap mapStashSt,
                     const float *__restrict__ pi, const float *__restrict__ Bsm, BwObs ob, int64_t B, int64_t T,
                     float *__restrict__ scale, double *__restrict__ loglik, long long *trace) {
    constexpr bool SPLIT = PREC != KHMM_TC_TF32;
    // (Dropping the W_lo A_hi product -- two MMAs, the recursion vector rounded to bf16 -- was measured: the aggregate
    // counts stay within 6e-6 but single emission counts scatter by up to 4e-3 at 4,096 sequences, and the kernels are
    // no faster than with tf32 operands; not kept.)
    constexpr int NPASS = 3;
#define FW_TRACE(ev)                                                                                      \
    do {                                                                                                  \
        if (trace && blockIdx.x == 0 && tile == blockIdx.x && t >= 8 && t < 16 && (threadIdx.x & 31) == 0) \
            trace[256 + (t - 8) * 32 + (ev)] = clock64();                                                 \
    } while (0)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // round up to 1024 B by pointer arithmetic on the shared array (an integer round trip would turn every
    // later access into a generic LD/ST that the compiler cannot reorder against global stores)
    unsigned char *tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *ebuf = tiles + BW_TILE;
    float *csm = reinterpret_cast<float *>(ebuf + 2 * FW_EBUF);   // [2][128] scale factors of the step in flight
    float *psm = csm + 256;                                       // [2 parities][2 halves][128] partial row sums
    FwdBars *bars = (FwdBars *)(psm + 512);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntiles = (B + 127) / 128;

    if (threadIdx.x == 0) {
        mbar_init(&bars->mat_full, 1);
        mbar_init(&bars->a_ready, 256);
        mbar_init(&bars->d_full[0], 1);
        mbar_init(&bars->d_full[1], 1);
        for (int k = 0; k < 2; k++) {
            mbar_init(&bars->e_full[k], 64);
            mbar_init(&bars->w_ready[k], 256);
            mbar_init(&bars->buf_free[k], 1);
        }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(&bars->tmem_base, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;   // columns [0,128): accumulator D, [128,256): A operand W

    if (warp == 8) {
        if (elect_one()) {
            tma_prefetch_desc(&mapAT);
            mbar_arrive_expect_tx(&bars->mat_full, BW_TILE);
            if (SPLIT) {
                // A^T_hi | A^T_lo: bf16 [128][128] each, two sub-tiles of [128 rows][64 K elements = 128 B]
                tma_prefetch_desc(&mapATlo);
                for (int kb = 0; kb < 2; kb++) {
                    tma_load_2d(tiles + kb * BW_SUB, &mapAT, &bars->mat_full, kb * 64, 0);
                    tma_load_2d(tiles + 2 * BW_SUB + kb * BW_SUB, &mapATlo, &bars->mat_full, kb * 64, 0);
                }
            } else {
                for (int kb = 0; kb < 4; kb++) tma_load_2d(tiles + kb * BW_SUB, &mapAT, &bars->mat_full, kb * 32, 0);
            }
            mbar_wait(&bars->mat_full, 0);
            // the MMA of a step is issued as two column halves (N = 64) with a commit each: the epilogue warps of columns
            // 0-63 start while the tensor core still works on columns 64-127
            const uint32_t idesc = SPLIT ? make_idesc_bf16(128, 64) : make_idesc_tf32(128, 64);
            uint32_t na = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int64_t t = 1; t < T; t++) {
                    mbar_wait(&bars->a_ready, na & 1);
                    na++;
                    FW_TRACE(0);
                    tc_fence_after();
#pragma unroll
                    for (int nh = 0; nh < 2; nh++) {
                        // output columns [64nh, 64nh + 64) = rows 64nh.. of the K-major A^T tile (8 KB into every sub-tile)
                        if (SPLIT) {
                            // (W_hi, A_hi), (W_hi, A_lo), (W_lo, A_hi): W_hi in TMEM columns [128,192), W_lo in [192,256)
#pragma unroll
                            for (int pass = 0; pass < NPASS; pass++) {
                                const uint32_t acol = tmem + 128 + (pass == 2 ? 64 : 0);
                                const unsigned char *bt = tiles + (pass == 1 ? 2 * BW_SUB : 0) + nh * 8192;
#pragma unroll
                                for (int k = 0; k < 8; k++) {
                                    uint64_t db = make_kmajor_sw128_desc(bt + (k >> 2) * BW_SUB);
                                    mma_bf16_ts(tmem + 64 * nh, acol + k * 8, desc_advance(db, (k & 3) * 32), idesc, (pass | k) != 0);
                                }
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 16; k++) {
                                uint64_t db = make_kmajor_sw128_desc(tiles + (k >> 2) * BW_SUB + nh * 8192);
                                mma_tf32_ts(tmem + 64 * nh, tmem + 128 + k * 8, desc_advance(db, (k & 3) * 32), idesc, k != 0);
                            }
                        }
                        tc_commit(&bars->d_full[nh]);
                    }
                    FW_TRACE(1);
                }
        }
    } else if (warp == 9 || warp == 10) {
        // ------------------------------------------------ gather: emission rows of step t -> buffer t&1
        const int rbase = (warp - 9) * 64;          // this warp's 64 rows
        const int rsub = lane >> 3, piece = lane & 7;
        uint32_t nuse[2] = {0, 0};
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int64_t b0 = tile * 128 + rbase + lane, b1 = b0 + 32;
            b0 = b0 < B ? b0 : B - 1;
            b1 = b1 < B ? b1 : B - 1;
            for (int64_t t = 0; t < T; t++) {
                const int buf = (int)(t & 1);
                const int s0 = ob.sym(b0, t), s1 = ob.sym(b1, t);
                mbar_wait_relaxed(&bars->buf_free[buf], (nuse[buf] & 1) ^ 1);      // a step of slack
                nuse[buf]++;
                if (warp == 9) FW_TRACE(4);
                unsigned char *dstb = ebuf + buf * FW_EBUF;
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const int rr = 4 * k + rsub;
                    const int sy = __shfl_sync(0xffffffffu, k < 8 ? s0 : s1, rr & 31);
                    const int row = rbase + rr;
                    const float *src = Bsm + (int64_t)sy * BW_N + piece * 4;
                    unsigned char *d = dstb + row * 128 + ((piece ^ (row & 7)) << 4);
#pragma unroll
                    for (int c = 0; c < 4; c++) cp_async16(d + c * BW_SUB, src + c * 32);
                }
                cp_async_arrive(&bars->e_full[buf]);
                if (warp == 9) FW_TRACE(5);
            }
        }
    } else if (warp == 11) {
        // ------------------------------------------------ storer: one thread, TMA tensor stores of the W_t tile
        if (elect_one()) {
            tma_prefetch_desc(&mapStashSt);
            uint32_t nuse[2] = {0, 0};
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int64_t t = 0; t < T; t++) {
                    const int buf = (int)(t & 1);
                    mbar_wait_relaxed(&bars->w_ready[buf], nuse[buf] & 1);      // off the MMA -> epilogue chain
                    nuse[buf]++;
                    FW_TRACE(8);
                    for (int c = 0; c < 4; c++)
                        tma_store_3d(&mapStashSt, ebuf + buf * FW_EBUF + c * BW_SUB, c * 32, 0, (int)(tile * T + t));
                    bulk_store(scale + (tile * T + t) * 128, csm + buf * 128, 512);
                    bulk_commit();
                    bulk_wait_read0();          // shared memory has been read: the gather warps may refill the buffer
                    mbar_arrive(&bars->buf_free[buf]);
                    FW_TRACE(9);
                }
            bulk_wait0();
        }
    } else if (warp < 8) {
        const int row = threadIdx.x & 127;
        const int half = warp >> 2;                  // column half: chunks {0,1} or {2,3}
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int sw7 = row & 7;
        uint32_t nd = 0, ne[2] = {0, 0};
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t b = tile * 128 + row;
            const bool live = b < B;
            const int len_b = live ? ob.len(b) : 0;   // steps of this sequence (0: a row past the batch)
            float r = 1.f;           // 1 / c_{t-1}
            float prod = 1.f;        // mantissa of prod_t c_t
            int esum = 0;            // its exponent
            for (int64_t t = 0; t < T; t++) {
                const int cur = (int)(t & 1);
                unsigned char *myrow = ebuf + cur * FW_EBUF + row * 128;
                if (t > 0) {
                    mbar_wait(&bars->d_full[half], nd & 1);
                    nd++;
                    tc_fence_after();
                }
                if (warp == 0) FW_TRACE(12);
                mbar_wait(&bars->e_full[cur], ne[cur] & 1);
                ne[cur]++;
                if (warp == 0) FW_TRACE(13);
                float rowsum = 0.f;
#pragma unroll 1
                for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 32) {
                    uint32_t acc[32];
                    float e[32];
                    if (t > 0) tmem_ld_32x32(tmem + lane_base + c0, acc);
                    unsigned char *erow = myrow + (c0 >> 5) * BW_SUB;
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        float4 f = *reinterpret_cast<const float4 *>(erow + ((q ^ sw7) << 4));
                        e[4 * q] = f.x; e[4 * q + 1] = f.y; e[4 * q + 2] = f.z; e[4 * q + 3] = f.w;
                    }
                    if (t > 0) {
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            float4 f = ldg4(pi + c0 + 4 * q);
                            acc[4 * q] = __float_as_uint(f.x); acc[4 * q + 1] = __float_as_uint(f.y);
                            acc[4 * q + 2] = __float_as_uint(f.z); acc[4 * q + 3] = __float_as_uint(f.w);
                        }
                    }
                    uint32_t w[32];
#pragma unroll
                    for (int q = 0; q < 32; q++) {
                        float v = __uint_as_float(acc[q]) * r * e[q];
                        rowsum += v;
                        e[q] = t < len_b ? v : 0.f;   // rows past the batch / steps past the sequence: zeros in the stash
                        w[q] = SPLIT ? __float_as_uint(v) : tf32_bits(v);
                    }
                    if (SPLIT) {
                        if (t + 1 < T) {
                            uint32_t hi[16], lo[16];
#pragma unroll
                            for (int m = 0; m < 16; m++)
                                split_bf16x2(__uint_as_float(w[2 * m]), __uint_as_float(w[2 * m + 1]), hi[m], lo[m]);
                            tmem_st_32x16(tmem + lane_base + 128 + c0 / 2, hi);
                            if (NPASS == 3) tmem_st_32x16(tmem + lane_base + 192 + c0 / 2, lo);
                        }
                    } else if (t + 1 < T) {
                        tmem_st_32x32(tmem + lane_base + 128 + c0, w);
                    }
#pragma unroll
                    for (int q = 0; q < 8; q++)
                        *reinterpret_cast<float4 *>(erow + ((q ^ sw7) << 4)) = make_float4(e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3]);
                }
                // the two column halves of a row exchange their partial sums (named barrier of the two warps)
                psm[(cur * 2 + half) * 128 + row] = rowsum;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
                rowsum += psm[(cur * 2 + (half ^ 1)) * 128 + row];
                if (half == 0) csm[cur * 128 + row] = rowsum;
                if (t + 1 < T) {
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&bars->a_ready);
                }
                fence_proxy_async();         // the tile leaves through the async proxy (TMA store)
                mbar_arrive(&bars->w_ready[cur]);
                if (warp == 0) FW_TRACE(14);
                r = 1.0f / rowsum;
                if (t < len_b) {     // the padding's scale factors are not part of log P
                    int ex;
                    prod = frexpf(prod * rowsum, &ex);
                    esum += ex;
                }
            }
            if (live && half == 0) loglik[b] = (double)esum * 0.6931471805599453 + log((double)prod);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------ backward
struct __align__(8) BwdBars {
    uint64_t mat_full, alpha_full, alpha_empty, v_ready, d1_full, mma2_done, stage_full, stage_empty, e_full[2], e_free[2];
    uint32_t tmem_base;
};
#ifndef KHMM_RED_DEFER
#define KHMM_RED_DEFER 0
#endif
#ifndef KHMM_RED_PACE_NS
#define KHMM_RED_PACE_NS 40
#endif
// measured at c4 size (backward kernel): no pacing 19.2-19.6 ms, 40-60 ns per reduction 16.3-16.8 ms, 80-120 ns 18.4 ms,
// 200 ns 32 ms (the reduction warps become the critical path); deferring the burst behind v_ready: no gain
constexpr bool RED_DEFER = KHMM_RED_DEFER;
constexpr int RED_PACE_NS = KHMM_RED_PACE_NS;
constexpr int BW_ECHUNK = 128 * 128;   // emission rows of one 32-column chunk: [128 rows][128 B], 16-byte pieces XOR (row & 7)

// Optional timeline (development): clock64 stamps of block 0, first tile, backward steps 8..15.
#define BW_TRACE(ev)                                                                                   \
    do {                                                                                               \
        if (out.trace && blockIdx.x == 0 && tile == blockIdx.x && (T - 1 - t) >= 8 && (T - 1 - t) < 16 && \
            (threadIdx.x & 31) == 0)                                                                   \
            out.trace[((T - 1 - t) - 8) * 32 + (ev)] = clock64();                                       \
    } while (0)

struct BwdOut {
    long long *trace;  // nullptr unless khmm_debug_bw_trace() armed it
    float *tab;        // [BW_REPL][M+2][128] fp32 count table (row M: pi, row M+1: gamma_ref at t = T-1)
    double *xi;        // counts.xi [128][128]
    double *flags;     // counts tail + 2
};

// Warp roles (384 threads): warps 0-3 epilogue (setmaxnreg.inc 232), warps 4-7 count reductions (168),
// warp 8 TMA producer, warp 9 MMA issuer + TMEM allocator, warps 10-11 idle (96): three warps share one
// SM sub-partition (16K registers), so the register file is re-split by role.
// SPLIT: MMA1 on bf16 hi/lo pairs like the forward kernel (V_hi, V_lo in TMEM, A_hi, A_lo in shared memory); MMA2 stays tf32
// (its rounding errors are unbiased and average out over the B*T rows it sums).
template <int PREC>
__global__ void __launch_bounds__(384, 1)
bw_backward_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
                      const __grid_constant__ CUtensorMap mapStash,
                      const float *__restrict__ Bsm, BwObs ob, int64_t B, int64_t T, const float *__restrict__ scale,
                      BwdOut out) {
    constexpr bool SPLIT = PREC != KHMM_TC_TF32;
    constexpr int NPASS = 3;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sA = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // A, K-major SW128
    unsigned char *sW = sA + BW_TILE;      // W_t tile, MN-major (TMA, SWIZZLE_128B_ATOM_32B)
    unsigned char *sV = sW + BW_TILE;      // V2_{t+1} tile, MN-major (written by the epilogue)
    unsigned char *sE = sV + BW_TILE;      // two emission chunk buffers filled by the gather warps
    BwdBars *bars = (BwdBars *)(sE + 2 * BW_ECHUNK);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntiles = (B + 127) / 128;
    const int M = ob.M;

    if (threadIdx.x == 0) {
        mbar_init(&bars->mat_full, 1);
        mbar_init(&bars->alpha_full, 1);
        mbar_init(&bars->alpha_empty, 5);
        mbar_init(&bars->v_ready, 128);
        mbar_init(&bars->d1_full, 1);
        mbar_init(&bars->mma2_done, 1);
        mbar_init(&bars->stage_full, 128);
        mbar_init(&bars->stage_empty, 4);
        for (int k = 0; k < 2; k++) {
            mbar_init(&bars->e_full[k], 64);
            mbar_init(&bars->e_free[k], 4);
        }
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp >= 10) {
        // ------------------------------------------------ gather warps: emission rows b(o_t), one 32-column chunk
        // at a time, into shared memory with coalesced cp.async (lane = 16-byte piece, 4 rows per instruction).
        // A thread-per-row gather touches 32 lines per load instruction and, together with the count
        // reductions, saturated the LSU (13,000 cycles per step in the first version).
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        const int rbase = (warp - 10) * 64;
        const int rsub = lane >> 3, piece = lane & 7;
        uint32_t nchunk = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int64_t b0 = tile * 128 + rbase + lane, b1 = b0 + 32;
            b0 = b0 < B ? b0 : B - 1;
            b1 = b1 < B ? b1 : B - 1;
            for (int64_t t = T - 1; t >= 1; t--) {   // step t = 0 forms no V_t: no emission rows needed
                const int s0 = ob.sym(b0, t), s1 = ob.sym(b1, t);
                for (int c = 0; c < 4; c++, nchunk++) {
                    const int buf = nchunk & 1;
                    mbar_wait(&bars->e_free[buf], ((nchunk >> 1) & 1) ^ 1);
                    unsigned char *dstb = sE + buf * BW_ECHUNK;
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        const int rr = 4 * k + rsub;                       // row within this warp's 64
                        const int sy = __shfl_sync(0xffffffffu, k < 8 ? s0 : s1, rr & 31);
                        const int row = rbase + rr;
                        cp_async16(dstb + row * 128 + ((piece ^ (row & 7)) << 4),
                                   Bsm + (int64_t)sy * BW_N + c * 32 + piece * 4);
                    }
                    cp_async_arrive(&bars->e_full[buf]);
                }
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------ TMA producer: A once, then W_t tiles of the stash
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        if (elect_one()) {
            tma_prefetch_desc(&mapA);
            tma_prefetch_desc(&mapStash);
            mbar_arrive_expect_tx(&bars->mat_full, BW_TILE);
            if (SPLIT) {
                tma_prefetch_desc(&mapAlo);
                for (int kb = 0; kb < 2; kb++) {
                    tma_load_2d(sA + kb * BW_SUB, &mapA, &bars->mat_full, kb * 64, 0);
                    tma_load_2d(sA + 2 * BW_SUB + kb * BW_SUB, &mapAlo, &bars->mat_full, kb * 64, 0);
                }
            } else {
                for (int kb = 0; kb < 4; kb++) tma_load_2d(sA + kb * BW_SUB, &mapA, &bars->mat_full, kb * 32, 0);
            }
            uint32_t np = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int64_t t = T - 1; t >= 0; t--) {
                    mbar_wait(&bars->alpha_empty, (np & 1) ^ 1);
                    np++;
                    BW_TRACE(0);
                    mbar_arrive_expect_tx(&bars->alpha_full, BW_TILE);
                    for (int c = 0; c < 4; c++)
                        tma_load_3d(sW + c * BW_SUB, &mapStash, &bars->alpha_full, c * 32, 0, (int)(tile * T + t));
                }
        }
    } else if (warp == 9) {
        // ------------------------------------------------ MMA issuer
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        if (elect_one()) {
            mbar_wait(&bars->mat_full, 0);
            const uint32_t idesc_k = SPLIT ? make_idesc_bf16(128, 128) : make_idesc_tf32(128, 128);
            const uint32_t idesc_mn = make_idesc_tf32(128, 128, true, true);
            const uint64_t dW = make_mnmajor_sw128a32_desc(sW, BW_SUB, 512);
            const uint64_t dV = make_mnmajor_sw128a32_desc(sV, BW_SUB, 512);
            uint32_t nv = 0, nfull = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                nfull++;   // step t = T-1 has no MMA2 (the epilogue releases W_{T-1} twice on our behalf)
                for (int64_t t = T - 2; t >= 0; t--) {
                    mbar_wait(&bars->v_ready, nv & 1);
                    nv++;
                    BW_TRACE(4);
                    tc_fence_after();
                    if (SPLIT) {
                        // (V_hi, A_hi), (V_hi, A_lo), (V_lo, A_hi): V_hi in TMEM columns [COL_V, +64), V_lo in [COL_V+64, +128)
#pragma unroll
                        for (int pass = 0; pass < NPASS; pass++) {
                            const uint32_t acol = tmem + COL_V + (pass == 2 ? 64 : 0);
                            const unsigned char *bt = sA + (pass == 1 ? 2 * BW_SUB : 0);
#pragma unroll
                            for (int k = 0; k < 8; k++) {
                                uint64_t db = make_kmajor_sw128_desc(bt + (k >> 2) * BW_SUB);
                                mma_bf16_ts(tmem + COL_D1, acol + k * 8, desc_advance(db, (k & 3) * 32), idesc_k, (pass | k) != 0);
                            }
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; k++) {
                            uint64_t db = make_kmajor_sw128_desc(sA + (k >> 2) * BW_SUB);
                            mma_tf32_ts(tmem + COL_D1, tmem + COL_V + k * 8, desc_advance(db, (k & 3) * 32), idesc_k, k != 0);
                        }
                    }
                    tc_commit(&bars->d1_full);
                    BW_TRACE(5);
                    mbar_wait(&bars->alpha_full, nfull & 1);
                    nfull++;
                    BW_TRACE(6);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 16; k++)
                        mma_tf32_ss(tmem + COL_XI, desc_advance(dW, k * 1024), desc_advance(dV, k * 1024), idesc_mn,
                                    (t != T - 2) || k != 0);
                    tc_commit(&bars->mma2_done);
                    tc_commit(&bars->alpha_empty);
                    BW_TRACE(7);
                }
            }
        }
    } else if (warp < 4) {
        // ------------------------------------------------ epilogue: thread = sequence = TMEM lane
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int row = threadIdx.x;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const int sw = row & 3;   // Swizzle<2,5,2>: 32-byte chunk index ^= row & 3
        const int hb = (row >> 2) & 1;
        uint32_t nd = 0, nm = 0, nfull = 0, nstage = 0, nchunk = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t b = tile * 128 + row;
            const bool live = b < B;
            const int len_b = live ? ob.len(b) : 0;
            for (int64_t t = T - 1; t >= 0; t--) {
                const bool last = (t == T - 1);
                const bool bone = last || t >= len_b - 1;      // beta_t = 1: the sequence's own last step, or its padding
                const float *sct = scale + (tile * T + t) * 128 + row;   // tile-major scale [tile][t][row]
                const float rc1 = last ? 1.f : 1.0f / sct[128];
                const float rcm1 = (t > 0 && t < len_b) ? 1.0f / sct[-128] : 0.f;   // V2_t = 0 in the padding
                if (!last) {
                    mbar_wait(&bars->d1_full, nd & 1);
                    nd++;
                    tc_fence_after();
                }
                if (warp == 0) BW_TRACE(8);
                mbar_wait(&bars->alpha_full, nfull & 1);
                nfull++;
                if (warp == 0) BW_TRACE(9);
                // gamma_ref chunk: W_t . beta_t  (beta = 1 at t = T-1)
                auto gchunk = [&](int c0, float (&g)[32], float (&beta)[32]) {
                    uint32_t acc[32];
                    if (!last) tmem_ld_32x32(tmem + lane_base + COL_D1 + c0, acc);
                    const unsigned char *wrow = sW + (c0 >> 5) * BW_SUB + row * 128;
                    // rows r and r+4 share a 32-byte slot: lanes with (row & 4) read the upper half first so that
                    // 8 consecutive lanes hit 8 distinct 16-byte bank groups, then the halves are swapped back
#pragma unroll
                    for (int c32 = 0; c32 < 4; c32++) {
                        const unsigned char *slot = wrow + ((c32 ^ sw) << 5);
                        float4 x = *reinterpret_cast<const float4 *>(slot + (hb << 4));
                        float4 y = *reinterpret_cast<const float4 *>(slot + ((hb ^ 1) << 4));
                        float4 lo = hb ? y : x, hi = hb ? x : y;
                        int q = c32 * 8;
                        g[q] = lo.x; g[q + 1] = lo.y; g[q + 2] = lo.z; g[q + 3] = lo.w;
                        g[q + 4] = hi.x; g[q + 5] = hi.y; g[q + 6] = hi.z; g[q + 7] = hi.w;
                    }
                    if (!last) tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; q++) {
                        beta[q] = bone ? 1.f : __uint_as_float(acc[q]) * rc1;
                        g[q] *= beta[q];
                    }
                };
                float sub = 0.f, inv = 1.f, floorv = 0.f;
                if (t == 0) {
                    // pi_new = smooth1d(gamma[0]) in place (discrete_hmm.py:132, util.py:56-60)
                    int k = 0;
#pragma unroll 1
                    for (int c0 = 0; c0 < BW_N; c0 += 32) {
                        float g[32], beta[32];
                        gchunk(c0, g, beta);
#pragma unroll
                        for (int q = 0; q < 32; q++) k += (g[q] < 1e-10f) ? 1 : 0;
                    }
                    if (k >= BW_N) {
                        if (live) atomicAdd(out.flags, 1.0);
                        k = BW_N - 1;
                    }
                    sub = (float)((double)k * 1e-10 / (double)(BW_N - k));
                    floorv = 1e-10f;
                    float sum = 0.f;
#pragma unroll 1
                    for (int c0 = 0; c0 < BW_N; c0 += 32) {
                        float g[32], beta[32];
                        gchunk(c0, g, beta);
#pragma unroll
                        for (int q = 0; q < 32; q++) sum += fmaxf(g[q] - sub, floorv);
                    }
                    inv = 1.0f / sum;
                }
                mbar_wait(&bars->stage_empty, (nstage & 1) ^ 1);
                nstage++;
                if (warp == 0) BW_TRACE(10);
                tc_fence_after();
#pragma unroll 1
                for (int c0 = 0; c0 < BW_N; c0 += 32) {
                    float g[32], beta[32];
                    gchunk(c0, g, beta);
                    uint32_t u[32];
                    // staging column 16*blk + 8*n2 + 2*j + e holds logical column 16*blk + 4*j + 2*n2 + e, so that the
                    // four registers a lane gets from two adjacent 16x256b repetitions are 4 CONSECUTIVE columns
                    // (one red.v4 per lane, a quad = 64 contiguous bytes of a row)
                    if (t == 0) {
#pragma unroll
                        for (int q = 0; q < 32; q++) {
                            const int lq = (q & 16) | ((q & 6) << 1) | ((q & 8) >> 2) | (q & 1);
                            u[q] = __float_as_uint(fmaxf(g[lq] - sub, floorv) * inv);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; q++) {
                            const int lq = (q & 16) | ((q & 6) << 1) | ((q & 8) >> 2) | (q & 1);
                            u[q] = __float_as_uint(g[lq]);
                        }
                    }
                    tmem_st_32x32(tmem + lane_base + COL_STG + c0, u);
                    if (t > 0) {
                        const int buf = nchunk & 1;
                        mbar_wait(&bars->e_full[buf], (nchunk >> 1) & 1);
                        nchunk++;
                        const unsigned char *erow = sE + buf * BW_ECHUNK + row * 128;
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            float4 f = *reinterpret_cast<const float4 *>(erow + ((q ^ (row & 7)) << 4));
                            if (SPLIT) {
                                split_bf16x2(f.x * beta[4 * q], f.y * beta[4 * q + 1], u[2 * q], u[16 + 2 * q]);
                                split_bf16x2(f.z * beta[4 * q + 2], f.w * beta[4 * q + 3], u[2 * q + 1], u[16 + 2 * q + 1]);
                            } else {
                                u[4 * q] = tf32_bits(f.x * beta[4 * q]);
                                u[4 * q + 1] = tf32_bits(f.y * beta[4 * q + 1]);
                                u[4 * q + 2] = tf32_bits(f.z * beta[4 * q + 2]);
                                u[4 * q + 3] = tf32_bits(f.w * beta[4 * q + 3]);
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->e_free[buf]);
                        if (SPLIT) {
                            // u[0..15] = hi pairs, u[16..31] = lo pairs of the 32 values of this chunk
                            uint32_t hi[16], lo[16];
#pragma unroll
                            for (int m = 0; m < 16; m++) { hi[m] = u[m]; lo[m] = u[16 + m]; }
                            tmem_st_32x16(tmem + lane_base + COL_V + c0 / 2, hi);
                            tmem_st_32x16(tmem + lane_base + COL_V + 64 + c0 / 2, lo);
                        } else {
                            tmem_st_32x32(tmem + lane_base + COL_V + c0, u);
                        }
                    }
                }
                // W_t has been read: hand the buffer back to the TMA producer (MMA2 adds its own arrival)
                if (warp == 0) BW_TRACE(11);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->alpha_empty);
                // t = T-1 has no MMA2: its arrival is made here, inside the same barrier phase (an early
                // arrival by the MMA thread could land in the previous tile's last phase)
                if (last && threadIdx.x == 0) mbar_arrive(&bars->alpha_empty);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars->stage_full);
                if (warp == 0) BW_TRACE(12);
                if (!last) {
                    // MMA2(t) reads V2_{t+1} from shared memory: wait before overwriting it
                    mbar_wait(&bars->mma2_done, nm & 1);
                    nm++;
                    tc_fence_after();
                }
                if (warp == 0) BW_TRACE(13);
                if (t > 0) {
                    // V2_t = V_t / c_{t-1} -> shared memory, MN-major Swizzle<2,5,2> (B operand of MMA2(t-1))
#pragma unroll 1
                    for (int c0 = 0; c0 < BW_N; c0 += 32) {
                        uint32_t v[32];
                        if (SPLIT) {
                            uint32_t hi[16], lo[16];
                            tmem_ld_32x16(tmem + lane_base + COL_V + c0 / 2, hi);
                            tmem_ld_32x16(tmem + lane_base + COL_V + 64 + c0 / 2, lo);
                            tmem_ld_wait();
#pragma unroll
                            for (int m = 0; m < 16; m++) {
                                v[2 * m] = tf32_bits(unsplit_lo(hi[m], lo[m]) * rcm1);
                                v[2 * m + 1] = tf32_bits(unsplit_hi(hi[m], lo[m]) * rcm1);
                            }
                        } else {
                            tmem_ld_32x32(tmem + lane_base + COL_V + c0, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int q = 0; q < 32; q++) v[q] = tf32_bits(__uint_as_float(v[q]) * rcm1);
                        }
                        unsigned char *vrow = sV + (c0 >> 5) * BW_SUB + row * 128;
#pragma unroll
                        for (int c32 = 0; c32 < 4; c32++) {
                            const int q = c32 * 8;
                            uint4 lo = make_uint4(v[q], v[q + 1], v[q + 2], v[q + 3]);
                            uint4 hi = make_uint4(v[q + 4], v[q + 5], v[q + 6], v[q + 7]);
                            unsigned char *slot = vrow + ((c32 ^ sw) << 5);
                            *reinterpret_cast<uint4 *>(slot + (hb << 4)) = hb ? hi : lo;          // same conflict-free order
                            *reinterpret_cast<uint4 *>(slot + ((hb ^ 1) << 4)) = hb ? lo : hi;
                        }
                    }
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive(&bars->v_ready);
                    if (warp == 0) BW_TRACE(14);
                } else if (T > 1) {
                    // tile done: Xi[i][j] (lane = i) -> float64 counts
#pragma unroll 1
                    for (int c0 = 0; c0 < BW_N; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32(tmem + lane_base + COL_XI + c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 32; q++) atomicAdd(out.xi + row * BW_N + c0 + q, (double)__uint_as_float(v[q]));
                    }
                    tc_fence_before();
                }
            }
        }
    } else if (warp < 8) {
        // ------------------------------------------------ RED warps: staging rows -> fp32 count table in L2
        const int quad = warp & 3;    // TMEM lane quadrant this warp may access
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        float *tab = out.tab + (size_t)(blockIdx.x % BW_REPL) * (size_t)(M + 2) * BW_N;
        const int colq = 4 * (lane & 3);   // a quad covers 16 consecutive columns with one red.v4 per lane
        uint32_t ns = 0, nvr = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int64_t bq[4];
            bool lv[4];
            int lenq[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                bq[k] = tile * 128 + quad * 32 + k * 8 + (lane >> 2);
                lv[k] = bq[k] < B;
                lenq[k] = lv[k] ? ob.len(bq[k]) : 0;
            }
            for (int64_t t = T - 1; t >= 0; t--) {
                int sym[4];
#pragma unroll
                for (int k = 0; k < 4; k++) sym[k] = lv[k] ? ob.sym(bq[k], t) : 0;
                mbar_wait(&bars->stage_full, ns & 1);
                ns++;
                if (warp == 4) BW_TRACE(16);
                tc_fence_after();
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    uint32_t r[4][16];
#pragma unroll
                    for (int c = 0; c < 4; c++)
                        tmem_ld_16x256_x4(tmem + lane_base + ((uint32_t)(half * 16) << 16) + COL_STG + c * 32, r[c]);
                    tmem_ld_wait();
                    if (half == 1) {
                        // the staging area has been read: hand it back before issuing the reductions
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->stage_empty);
                        if (warp == 4) BW_TRACE(17);
                    }
                    if (RED_DEFER && half == 0 && t > 0) {
                        // LSU scheduling: the reductions (16 cycles of LSU time each) would otherwise sit in front of
                        // the epilogue's V2 stores; let them run under the next step's MMA instead
                        mbar_wait(&bars->v_ready, nvr & 1);
                        nvr++;
                    }
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int k = half * 2 + e;   // row = quad*32 + 8k + lane/4
                        if (t >= lenq[k]) continue;       // a row past the batch, or a step of the row's padding
                        float *dst = tab + (size_t)sym[k] * BW_N + colq;
#pragma unroll
                        for (int c = 0; c < 4; c++) {
#pragma unroll
                            for (int n = 0; n < 4; n += 2) {
                                red_add_v4(dst + c * 32 + 8 * n, __uint_as_float(r[c][4 * n + 2 * e]),
                                           __uint_as_float(r[c][4 * n + 2 * e + 1]), __uint_as_float(r[c][4 * n + 4 + 2 * e]),
                                           __uint_as_float(r[c][4 * n + 4 + 2 * e + 1]));
                                // optional pacing: a burst of reductions fills the LSU queue and delays the epilogue's
                                // shared-memory traffic and the MMA operand fetches behind it
                                if (RED_PACE_NS) asm volatile("nanosleep.u32 %0;" ::"n"(RED_PACE_NS));
                            }
                        }
                        if (t == 0 || t == lenq[k] - 1) {
#pragma unroll 1
                            for (int which = 0; which < 2; which++) {
                                if ((which == 0 && t != 0) || (which == 1 && t != lenq[k] - 1)) continue;
                                float *d2 = tab + (size_t)(M + which) * BW_N + colq;
#pragma unroll
                                for (int c = 0; c < 4; c++)
#pragma unroll
                                    for (int n = 0; n < 4; n += 2)
                                        red_add_v4(d2 + c * 32 + 8 * n, __uint_as_float(r[c][4 * n + 2 * e]),
                                                   __uint_as_float(r[c][4 * n + 2 * e + 1]),
                                                   __uint_as_float(r[c][4 * n + 4 + 2 * e]),
                                                   __uint_as_float(r[c][4 * n + 4 + 2 * e + 1]));
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem, 512);
}

// No "memory" clobber: nothing in the kernel reads the count table, and the clobber would pin every shared-memory access
// between two reductions.
__device__ __forceinline__ void red_add_v4_nc(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d));
}

#define BW2_TRACE(ev)                                                                                  \
    do {                                                                                               \
        if (out.trace && blockIdx.x == 0 && tile == blockIdx.x && (Ti - 1 - t) >= 8 && (Ti - 1 - t) < 16 && \
            (threadIdx.x & 31) == 0)                                                                   \
            out.trace[((Ti - 1 - t) - 8) * 32 + (ev)] = clock64();                                      \
    } while (0)

// ------------------------------------------------- backward, transposed, four quarter-tile pipelines in flight (round 2)
// Both kernels above run ONE tile per CTA, and a step of a tile is a serial chain -- MMA1 -> thread work -> MMA1 --
// whose links use different resources: their step period (8,100 / 8,300 cycles) is the SUM of the links.  This kernel keeps
// the transposed formulation (TMEM lane = state: every thread access is contiguous across the warp, the transition matrix
// is a resident TMEM operand) and splits the tile into two HALVES of 64 sequences that run as independent pipelines:
//     MMA1(t, h)  D1^T[i][b] = sum_j A[i][j] V_{t+1}[b][j]      M = 128 states, N = 64 sequences: a transposed MMA loses
//                                                               nothing at N = 64 (a thread-per-sequence M = 64 MMA would
//                                                               run at half rate)
//     MMA2(t, h)  Xi^T[j][i] += sum_b V2^T_{t+1}[j][b] W_t[b][i]  K = the 64 sequences of the half, one accumulator
// so the tensor pipe, the compute warps of the other half, the TMA stream and the L2 reductions overlap.  What the threads do
// per (sequence b, state i = thread), all warp-contiguous:
//     e   = Bsm[o_t(b)][i]            coalesced 128-byte read of a table row through L1/L2, prefetched one step ahead
//     w   = W_t[b][i]                 shared memory (the TMA tile, read in place; two buffers per half)
//     beta = D1^T[i][b] / c_{t+1}(b);  gamma_ref = w beta  -> red.global.add.f32 into row o_t(b) of the count table:
//                                     one coalesced 128-byte reduction per warp instruction, straight from registers
//     V   = e beta -> shared memory (K-major operand of MMA1(t-1)),  V2 = V / c_{t-1}(b) -> TMEM (operand of MMA2(t-1))
// SIXTEEN compute warps, warp (q, h, c) = states 32q..32q+31, sequences 32c..32c+31 of half h: four warps per scheduler (the
// eight-warp version ran each warp at 0.25 instructions per clock on its own dependency chains -- 5,000 cycles per
// half-step whatever was switched off).  The loop body is straight-line: unconditional reductions (a dead (sequence, step)
// pair adds the zero of its zero stash row), the rows of a sequence's last step summed in a register.
// No staging of gamma or of the emission rows in shared memory.  Shared memory: W 2 halves x 2 buffers x 32 KB, V 2 x 32 KB.
// TMEM: A (or A_hi | A_lo), D1^T (2 x 64 columns), Xi^T, V2^T (2 x 64 columns) = 512 columns.
// SPLIT (bf16x3): MMA1 = A_hi V_hi + A_lo V_hi + A_hi V_lo on bf16 hi/lo pairs (kind::f16), A_hi/A_lo resident in TMEM, V_hi/V_lo
// in the same 32 KB; MMA2 stays tf32 (its rounding errors are unbiased and average out over the B*T rows it sums).
constexpr uint32_t C3_A = 0, C3_R0 = 128, C3_XI = 256, C3_R1 = 384;   // R0 / R1: D1^T of one step, V2^T of the other (ping-pong)
constexpr int BW_HSUB = 64 * 128;       // one [64 sequences][128 B] sub-tile of a half (8 KB)
constexpr int BW_HTILE = 4 * BW_HSUB;   // one half: [64 sequences][128 states] fp32 (32 KB)
constexpr int BW3_PREP = 4;             // prep ring depth: the prep warps run up to three steps ahead of the compute warps
constexpr int BW3_THREADS = 640;        // 16 compute warps + TMA + MMA + 2 prep warps

struct __align__(8) Bwd3Bars {
    uint64_t w_full[2][2], w_empty[2][2], v_ready[4], d1_full[4], prep_full[2][BW3_PREP], prep_empty[2][BW3_PREP],
        xi_done, xi_free;
    uint32_t tmem_base;
};
struct __align__(16) Bwd3Prep {   // per-sequence scalars of one step of one half
    uint32_t roff[64]; // 128 * clamped symbol: float offset of the table row o_t(b)
    float mulb[64];    // c_t / c_{t+1}, or 0 where beta_t = 1 (t >= len - 1: the sequence's last step or its padding):
                       // MMA1 runs on V2_{t+1} = V_{t+1} / c_t, so D1 = (A V_{t+1}) / c_t and beta_t = D1 c_t / c_{t+1}
    float rcm1[64];    // 1 / c_{t-1} for 0 < t < len, else 0 (V2_t = 0: the pair (t-1, t) does not count)
    float lastf[64];   // 1 on the sequence's last step (t = len - 1), else 0
    int flag[64];      // bit 0: (sequence, step) not live (a row past the batch, a step of the padding); bit 1: last step
    int anylast;       // some sequence of the half has its last step here (the compute warps then read lastf)
    int pad[3];
};
// base + off floats with ONE 32-bit add: the host guarantees that neither table straddles a 4 GB boundary, so the high word
// of the address is a constant of the thread (the compiler's 64-bit LEA / LEA.HI.X pair per address was 4 of the 17.6
// instructions per (sequence, state) of the loop body)
struct Base32 {
    uint32_t lo, hi;
    __device__ __forceinline__ explicit Base32(const float *p) {
        const uint64_t a = reinterpret_cast<uint64_t>(p);
        lo = (uint32_t)a;
        hi = (uint32_t)(a >> 32);
    }
    __device__ __forceinline__ float *at(uint32_t off) const {
        uint64_t a;
        const uint32_t l = lo + (off << 2);
        asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(l), "r"(hi));
        return reinterpret_cast<float *>(a);
    }
};
__device__ __forceinline__ void red_add_f32_nc(float *p, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a));
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
static bool within_4gb_window(const void *p, size_t bytes) {
    const uintptr_t a = (uintptr_t)p;
    return bytes == 0 || (a >> 32) == ((a + bytes - 1) >> 32);
}

template <int PREC>
__global__ void __launch_bounds__(BW3_THREADS, 1)
bw_backward_tc3_kernel(const __grid_constant__ CUtensorMap mapStashH, const float *__restrict__ Amat,
                       const float *__restrict__ Bsm, BwObs ob, int64_t B, int64_t T, const float *__restrict__ scale,
                       BwdOut out) {
    constexpr bool SPLIT = PREC != KHMM_TC_TF32;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sW = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // [half][buffer][BW_HTILE]
    unsigned char *sV = sW + 4 * BW_HTILE;                                              // [half][BW_HTILE]
    Bwd3Prep *prep = (Bwd3Prep *)(sV + 2 * BW_HTILE);                                   // [half][BW3_PREP]
    Bwd3Bars *bars = (Bwd3Bars *)(prep + 2 * BW3_PREP);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntiles = (B + 127) / 128;
    const int M = ob.M;
    const int Ti = (int)T;
    const int64_t my_tiles = (int64_t)blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const uint32_t gtotal = (uint32_t)(my_tiles * Ti);     // steps of this CTA (all its tiles)

    if (threadIdx.x == 0) {
        for (int h = 0; h < 2; h++) {
            for (int k = 0; k < 2; k++) {
                mbar_init(&bars->w_full[h][k], 1);
                mbar_init(&bars->w_empty[h][k], 10);     // 8 compute warps + the commits of the two quarters' MMA2
            }
            for (int k = 0; k < BW3_PREP; k++) {
                mbar_init(&bars->prep_full[h][k], 1);
                mbar_init(&bars->prep_empty[h][k], 8);
            }
        }
        for (int pq = 0; pq < 4; pq++) {
            mbar_init(&bars->v_ready[pq], 128);
            mbar_init(&bars->d1_full[pq], 1);
        }
        mbar_init(&bars->xi_done, 1);
        mbar_init(&bars->xi_free, 256);
        fence_barrier_init();
    }
    if (warp == 17) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    if (warp < 8) {
        // the transition matrix, row i -> TMEM lane i, resident for the whole kernel: warp (q, hh) stores columns [64hh, 64hh+64)
        const int q = warp & 3, hh = warp >> 2;
        const int i = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        if (SPLIT) {
            // A_hi -> columns [0, 64), A_lo -> [64, 128): two consecutive K elements per 32-bit column
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int p = 0; p < 16; p++) {
                float4 f = ldg4(Amat + (size_t)i * BW_N + 64 * hh + 4 * p);
                split_bf16x2(f.x, f.y, hi[2 * p], lo[2 * p]);
                split_bf16x2(f.z, f.w, hi[2 * p + 1], lo[2 * p + 1]);
            }
            tmem_st_32x32(tmem + lane_base + C3_A + 32 * hh, hi);
            tmem_st_32x32(tmem + lane_base + C3_A + 64 + 32 * hh, lo);
        } else {
#pragma unroll 1
            for (int c0 = 64 * hh; c0 < 64 * hh + 64; c0 += 32) {
                uint32_t a[32];
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    float4 f = ldg4(Amat + (size_t)i * BW_N + c0 + 4 * p);
                    a[4 * p] = __float_as_uint(f.x); a[4 * p + 1] = __float_as_uint(f.y);
                    a[4 * p + 2] = __float_as_uint(f.z); a[4 * p + 3] = __float_as_uint(f.w);
                }
                tmem_st_32x32(tmem + lane_base + C3_A + c0, a);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 16) {
        // ------------------------------------------------ TMA producer: W_t half-tiles of the stash, two buffers per half
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (elect_one()) {
            tma_prefetch_desc(&mapStashH);
            uint32_t g = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int t = Ti - 1; t >= 0; t--, g++) {
                    const int buf = g & 1;
                    const uint32_t par = ((g >> 1) & 1) ^ 1;
                    for (int h = 0; h < 2; h++) {
                        mbar_wait_relaxed(&bars->w_empty[h][buf], par);      // a step of slack
                        if (h == 0) BW2_TRACE(0);
                        mbar_arrive_expect_tx(&bars->w_full[h][buf], BW_HTILE);
                        unsigned char *dst = sW + (h * 2 + buf) * BW_HTILE;
                        for (int c = 0; c < 4; c++)
                            tma_load_3d(dst + c * BW_HSUB, &mapStashH, &bars->w_full[h][buf], c * 32, 64 * h, (int)(tile * T + t));
                    }
                }
        }
    } else if (warp == 17) {
        // ------------------------------------------------ MMA issuer: the two halves alternate
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (elect_one()) {
            const uint32_t idesc1 = SPLIT ? make_idesc_bf16(128, 32) : make_idesc_tf32(128, 32);
            const uint32_t idesc2 = make_idesc_tf32(128, 128, false, true);
            uint32_t nv = 0, g = 0, ntile = 0;
            // operand descriptors: one base per half / per (half, buffer); a K step is a constant added to the address field
            // (no dynamically indexed arrays here: they would live in local memory, whose loads miss L1 in this kernel)
            const uint64_t dV0 = make_kmajor_sw128_desc(sV), dV1 = make_kmajor_sw128_desc(sV + BW_HTILE);
            const uint64_t dW0 = make_mnmajor_sw128a32_desc(sW, BW_HSUB, 512),
                           dW1 = make_mnmajor_sw128a32_desc(sW + 2 * BW_HTILE, BW_HSUB, 512);
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ntile++) {
                g++;      // step t = T-1 has no MMA
                for (int t = Ti - 2; t >= 0; t--, g++, nv++) {
                    const int buf = g & 1;
                    const uint32_t wpar = (g >> 1) & 1;
#pragma unroll
                    for (int pq = 0; pq < 4; pq++) {
                        // pipeline pq = 2h + c: the 32 sequences [32c, 32c + 32) of half h -- rows 32c.. of the half's V and W
                        // tiles (4 KB further into every sub-tile), columns 32c.. of its D1^T and V2^T
                        const int h = pq >> 1, c = pq & 1;
                        mbar_wait(&bars->v_ready[pq], nv & 1);
                        BW2_TRACE(pq == 0 ? 4 : (pq == 2 ? 20 : 31));
                        tc_fence_after();
                        const uint32_t dcol = tmem + ((g & 1) ? C3_R1 : C3_R0) + 64 * h + 32 * c;     // D1^T of this step
                        const uint32_t vcol = tmem + ((g & 1) ? C3_R0 : C3_R1) + 64 * h + 32 * c;     // V2^T of the previous step
                        // opaque to the optimiser: it would otherwise hoist all 80 per-instruction descriptors out of the time
                        // loop, spill them, and put local-memory loads (L2 latency here) in front of every MMA
                        uint64_t dvb = h ? dV1 : dV0, dwb = desc_advance(h ? dW1 : dW0, buf * BW_HTILE);
                        asm volatile("" : "+l"(dvb), "+l"(dwb));
                        // MMA1(t, pq): D1^T = A . V_{t+1}^T
                        if (SPLIT) {
                            // (A_hi, V_hi), (A_lo, V_hi), (A_hi, V_lo)
#pragma unroll
                            for (int pass = 0; pass < 3; pass++) {
                                const uint32_t acol = tmem + C3_A + (pass == 1 ? 64 : 0);
#pragma unroll
                                for (int k = 0; k < 8; k++)
                                    mma_bf16_ts(dcol, acol + k * 8,
                                                desc_advance(dvb, 4096 * c + (pass == 2 ? 2 * BW_HSUB : 0) + (k >> 2) * BW_HSUB + (k & 3) * 32),
                                                idesc1, (pass | k) != 0);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 16; k++)
                                mma_tf32_ts(dcol, tmem + C3_A + k * 8,
                                            desc_advance(dvb, 4096 * c + (k >> 2) * BW_HSUB + (k & 3) * 32), idesc1, k != 0);
                        }
                        tc_commit(&bars->d1_full[pq]);
                        BW2_TRACE(pq == 0 ? 5 : (pq == 2 ? 21 : 31));
                        if (c == 0) mbar_wait(&bars->w_full[h][buf], wpar);
                        // the previous tile's Xi must have been flushed before this tile's first MMA2 overwrites it
                        if (t == Ti - 2 && pq == 0 && ntile > 0) mbar_wait(&bars->xi_free, (ntile - 1) & 1);
                        BW2_TRACE(pq == 0 ? 6 : (pq == 2 ? 22 : 31));
                        tc_fence_after();
                        // MMA2(t, pq): Xi^T += V2^T_{t+1} . W_t over the 32 sequences of the quarter
                        const uint64_t dW = dwb;
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            mma_tf32_ts(tmem + C3_XI, vcol + k * 8, desc_advance(dW, 4096 * c + k * 1024), idesc2,
                                        (t != Ti - 2) || pq != 0 || k != 0);
                        tc_commit(&bars->w_empty[h][buf]);
                        if (t == 0 && pq == 3) tc_commit(&bars->xi_done);
                        BW2_TRACE(pq == 0 ? 7 : (pq == 2 ? 23 : 31));
                    }
                }
            }
        }
    } else if (warp >= 18) {
        // ------------------------------------------------ prep (one warp per half): per-sequence scalars, up to three steps ahead.
        // The symbol and scale loads of step g+1 are in flight while step g is written.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        const int h = warp - 18;
        int64_t bc[2] = {0, 0};
        int len[2] = {0, 0};
        int syn[2] = {0, 0}, lenn[2] = {0, 0};
        float c1n[2] = {1.f, 1.f}, c0n[2] = {1.f, 1.f}, cmn[2] = {1.f, 1.f};
        auto issue = [&](uint32_t gg) {      // the loads of step gg
            const int64_t tile = blockIdx.x + (int64_t)(gg / Ti) * gridDim.x;
            const int t = Ti - 1 - (int)(gg % Ti);
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int b = 64 * h + lane + 32 * k;
                if (t == Ti - 1) {      // first step of a tile: this lane's sequences change
                    const int64_t gb = tile * 128 + b;
                    bc[k] = gb < B ? gb : B - 1;
                    len[k] = gb < B ? ob.len(gb) : 0;
                }
                syn[k] = ob.sym(bc[k], t);
                c1n[k] = (t + 1 < Ti) ? scale[(tile * T + t + 1) * 128 + b] : 1.f;
                c0n[k] = scale[(tile * T + t) * 128 + b];
                cmn[k] = (t > 0) ? scale[(tile * T + t - 1) * 128 + b] : 1.f;
                lenn[k] = len[k];
            }
        };
        if (gtotal) issue(0);
        for (uint32_t g = 0; g < gtotal; g++) {
            const int t = Ti - 1 - (int)(g % Ti);
            int sy[2], lenc[2];
            float c1[2], c0[2], cm[2];
#pragma unroll
            for (int k = 0; k < 2; k++) { sy[k] = syn[k]; c1[k] = c1n[k]; c0[k] = c0n[k]; cm[k] = cmn[k]; lenc[k] = lenn[k]; }
            if (g + 1 < gtotal) issue(g + 1);
            const int buf = g % BW3_PREP;
            mbar_wait_relaxed(&bars->prep_empty[h][buf], ((g / BW3_PREP) & 1) ^ 1);     // three steps of slack
            Bwd3Prep &P = prep[h * BW3_PREP + buf];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int b = lane + 32 * k;
                const bool bone = (t == Ti - 1) || (t >= lenc[k] - 1);
                P.roff[b] = (uint32_t)(sy[k] * BW_N);
                P.flag[b] = (t < lenc[k] ? 0 : 1) | (t == lenc[k] - 1 ? 2 : 0);
                P.mulb[b] = bone ? 0.f : c0[k] / c1[k];
                P.rcm1[b] = (t > 0 && t < lenc[k]) ? 1.0f / cm[k] : 0.f;
                P.lastf[b] = (t == lenc[k] - 1) ? 1.f : 0.f;
            }
            {
                const unsigned any = __ballot_sync(0xffffffffu, (t == lenc[0] - 1) || (t == lenc[1] - 1));
                if (lane == 0) P.anylast = any != 0u;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->prep_full[h][buf]);
        }
    } else {
        // ------------------------------------------------ compute warps (q, h, c): thread = state i, 32 sequences
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        const int q = warp & 3, h = (warp >> 2) & 1, c = warp >> 3;
        const int pq = 2 * h + c;       // this warp's pipeline
        const int i = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        // byte offset of element (b, i) inside a half (b = sequence within the half, b = 32c + k):
        //   W, MN-major 32-byte atoms:            b * 128 + (mbase ^ ((b & 3) << 5))
        //   V tf32, K-major SWIZZLE_128B:         b * 128 + (kbase ^ ((b & 7) << 4))
        //   V bf16 pair (i even, i + 1), K-major: b * 128 + ((c16 ^ (b & 7)) << 4) + hbase   (+ 2 * BW_HSUB for the lo half)
        const uint32_t mbase = q * BW_HSUB + ((lane >> 3) << 5) + ((lane & 7) << 2);
        const uint32_t kbase = q * BW_HSUB + ((lane >> 2) << 4) + ((lane & 3) << 2);
        const uint32_t c16 = 4 * (q & 1) + (lane >> 3);
        const uint32_t hbase = (q >> 1) * BW_HSUB + 2 * (lane & 6);
        unsigned char *vk_c = sV + h * BW_HTILE + 32 * c * 128;       // this warp's 32 sequences of the V tile
        float *tabr = out.tab + (size_t)(blockIdx.x % BW_REPL) * (size_t)(M + 2) * BW_N;
        float *tab0 = tabr + i;
        const Base32 tabB(tab0), emB(Bsm + i);
        float e[32];
        if (gtotal) {
            mbar_wait(&bars->prep_full[h][0], 0);
            const Bwd3Prep &P0 = prep[h * BW3_PREP];
#pragma unroll
            for (int k4 = 0; k4 < 8; k4++) {
                const uint4 s4 = *reinterpret_cast<const uint4 *>(P0.roff + 32 * c + 4 * k4);
                e[4 * k4] = __ldg(emB.at(s4.x));
                e[4 * k4 + 1] = __ldg(emB.at(s4.y));
                e[4 * k4 + 2] = __ldg(emB.at(s4.z));
                e[4 * k4 + 3] = __ldg(emB.at(s4.w));
            }
        }
        uint32_t nd = 0, g = 0, ntile = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ntile++) {
            float glast = 0.f;     // gamma_ref[i] summed over the (sequence, last step) pairs of this tile with t > 0
            for (int t = Ti - 1; t >= 0; t--, g++) {
                const bool has_mma = (t <= Ti - 2);
                const int buf = g & 1;
                const int pb = g % BW3_PREP, pbn = (g + 1) % BW3_PREP;
                const Bwd3Prep &P = prep[h * BW3_PREP + pb];
                const Bwd3Prep &Pn = prep[h * BW3_PREP + pbn];
                const bool has_next = g + 1 < gtotal;
                mbar_wait(&bars->prep_full[h][pb], (g / BW3_PREP) & 1);
                if (warp == 0) BW2_TRACE(8);
                if (has_mma) {
                    mbar_wait(&bars->d1_full[pq], nd & 1);
                    nd++;
                    tc_fence_after();
                }
                if (warp == 0) BW2_TRACE(10);
                if (warp == 4) BW2_TRACE(24);
                mbar_wait(&bars->w_full[h][buf], (g >> 1) & 1);
                if (warp == 0) BW2_TRACE(9);
                if (has_next) mbar_wait(&bars->prep_full[h][pbn], ((g + 1) / BW3_PREP) & 1);   // long complete: three steps of slack
                if (warp == 0) BW2_TRACE(11);
                const unsigned char *w_c = sW + (h * 2 + buf) * BW_HTILE + 32 * c * 128;
                // D1^T of this step comes out of region R[g & 1] and V2^T of this step goes back IN PLACE, four columns at a time;
                // MMA2 of this step reads the other region (the previous step's V2^T), so no write-after-read wait is needed
                const uint32_t t_d1 = tmem + lane_base + ((g & 1) ? C3_R1 : C3_R0) + 64 * h + 32 * c;
                // Straight-line body over 8 groups of four sequences.  D1 comes out of TMEM and V2 goes back four columns at
                // a time: holding all 32 of each in registers spilled (ncu: the local-memory reloads missed L1 -- shared memory
                // takes it all -- and their L2 latency sat on every warp's dependency chain, 7,000 cycles per step).
                auto body = [&](auto t0c, auto mmac, auto lastc) {
                    constexpr bool T0 = decltype(t0c)::value;
                    constexpr bool MMA = decltype(mmac)::value;
                    constexpr bool LAST = decltype(lastc)::value;     // some sequence of the half ends at this step
                    float vprev[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int k4 = 0; k4 < 8; k4++) {
                        const int bq = 32 * c + 4 * k4;
                        uint32_t d4[4] = {0u, 0u, 0u, 0u};      // t = T-1: beta = 0 * 0 + 1
                        if (MMA) tmem_ld_32x4(t_d1 + 4 * k4, d4);
                        const float4 m4 = *reinterpret_cast<const float4 *>(P.mulb + bq);
                        const float4 r4 = *reinterpret_cast<const float4 *>(P.rcm1 + bq);
                        float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (LAST) l4 = *reinterpret_cast<const float4 *>(P.lastf + bq);
                        const uint4 s4 = *reinterpret_cast<const uint4 *>(P.roff + bq);
                        const float mu4[4] = {m4.x, m4.y, m4.z, m4.w};
                        const float rc4[4] = {r4.x, r4.y, r4.z, r4.w};
                        const float lf4[4] = {l4.x, l4.y, l4.z, l4.w};
                        const uint32_t sy4[4] = {s4.x, s4.y, s4.z, s4.w};
                        float w4[4], v4[4];
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) {
                            const int kb = 4 * k4 + kk;       // sequence within this warp's 32 (kb & 7 == b & 7, kb & 3 == b & 3)
                            w4[kk] = *reinterpret_cast<const float *>(w_c + kb * 128 + (mbase ^ ((kb & 3) << 5)));
                        }
                        if (MMA) tmem_ld_wait();
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) {
                            const int kb = 4 * k4 + kk;
                            // beta = D1 / c_{t+1}, or 1 on the sequence's last step and in its padding
                            const float beta = fmaf(__uint_as_float(d4[kk]), mu4[kk], mu4[kk] == 0.f ? 1.f : 0.f);
                            const float gm = w4[kk] * beta;       // gamma_ref: 0 for dead (sequence, step) pairs (zero stash row)
                            // V2_t = b(o_t) beta_t / c_{t-1}: the operand of MMA1(t-1) (shared memory) AND of MMA2(t-1) (TMEM)
                            const float v = (e[kb] * beta) * rc4[kk];
                            v4[kk] = v;
                            d4[kk] = tf32_bits(v);
                            if (!T0) {
                                red_add_f32_nc(tabB.at(sy4[kk]), gm);
                                if (LAST) glast = fmaf(gm, lf4[kk], glast);
                                if (!SPLIT) *reinterpret_cast<uint32_t *>(vk_c + kb * 128 + (kbase ^ ((kb & 7) << 4))) = d4[kk];
                            } else {
                                // t = 0: gamma_ref goes where V would (this quarter's own rows of the tile, fp32, the
                                // tf32 tile's addressing) for the smoothing pass below
                                *reinterpret_cast<float *>(vk_c + kb * 128 + (kbase ^ ((kb & 7) << 4))) = gm;
                            }
                        }
                        if (!T0) tmem_st_32x4(t_d1 + 4 * k4, d4);
                        if (SPLIT && !T0) {
                            // bf16 pairs: even lanes take the state pair (i, i+1) of sequence b, odd lanes the pair (i-1, i) of
                            // sequence b+4 -- rows b and b+4 put their 64 bytes into DIFFERENT halves of the 32 banks (rows b and
                            // b+1 share a half: ncu counted 1.96 wavefronts per store) -- so a group of four waits for its partner
                            // group: the stores of groups 2m and 2m+1 are issued together
                            if (k4 & 1) {
#pragma unroll
                                for (int kk = 0; kk < 4; kk++) {
                                    const int kb = 4 * (k4 - 1) + kk;          // partner: kb + 4
                                    const float mine = (lane & 1) ? vprev[kk] : v4[kk];
                                    const float other = __shfl_xor_sync(0xffffffffu, mine, 1);
                                    const float p0 = (lane & 1) ? other : vprev[kk];
                                    const float p1 = (lane & 1) ? v4[kk] : other;
                                    uint32_t hi, lo;
                                    split_bf16x2(p0, p1, hi, lo);
                                    const uint32_t offA = kb * 128 + ((c16 ^ (kb & 7)) << 4);
                                    const uint32_t offB = (kb + 4) * 128 + ((c16 ^ ((kb + 4) & 7)) << 4);
                                    unsigned char *dst = vk_c + hbase + ((lane & 1) ? offB : offA);
                                    *reinterpret_cast<uint32_t *>(dst) = hi;
                                    *reinterpret_cast<uint32_t *>(dst + 2 * BW_HSUB) = lo;
                                }
                            } else {
#pragma unroll
                                for (int kk = 0; kk < 4; kk++) vprev[kk] = v4[kk];
                            }
                        }
                        // the emission values of the next step (the next tile's first step after t = 0), these 4 sequences
                        if (has_next) {
                            const uint4 n4 = *reinterpret_cast<const uint4 *>(Pn.roff + bq);
                            e[4 * k4] = __ldg(emB.at(n4.x));
                            e[4 * k4 + 1] = __ldg(emB.at(n4.y));
                            e[4 * k4 + 2] = __ldg(emB.at(n4.z));
                            e[4 * k4 + 3] = __ldg(emB.at(n4.w));
                        }
                    }
                };
                const bool t0 = (t == 0);
                if (!t0) {
                    if (P.anylast) {
                        if (has_mma) body(std::false_type{}, std::true_type{}, std::true_type{});
                        else body(std::false_type{}, std::false_type{}, std::true_type{});
                    } else {
                        body(std::false_type{}, std::true_type{}, std::false_type{});    // no last step here implies t < T-1
                    }
                } else {
                    if (has_mma) body(std::true_type{}, std::true_type{}, std::false_type{});
                    else body(std::true_type{}, std::false_type{}, std::false_type{});
                }
                if (warp == 0) BW2_TRACE(14);
                // W_t has been read: hand the buffer back (t = T-1 has no MMA2: its arrival is made here as well)
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&bars->w_empty[h][buf]);
                    if (!has_mma && q == 0) mbar_arrive(&bars->w_empty[h][buf]);
                }
                if (!t0) {
                    tmem_st_wait();
                    fence_proxy_async();          // V is read by the tensor core through the async proxy
                    tc_fence_before();
                    mbar_arrive(&bars->v_ready[pq]);
                    if (warp == 0) BW2_TRACE(12);
                    if (warp == 4) BW2_TRACE(26);
                } else {
                    // t = 0: pi_new = smooth1d(gamma[0]) per sequence (util.py:56-60); the smoothed row counts for pi AND for
                    // the emission numerators (the reference overwrites gamma[0] in place, discrete_hmm.py:132).  The gamma
                    // rows of the half sit in its V buffer (no V_0 is formed); warp (q, c) takes 8 sequences.
                    named_bar_sync(1 + h, 256);
                    const unsigned char *scr = sV + h * BW_HTILE + (lane >> 3) * BW_HSUB;     // states 4 lane .. 4 lane + 3
#pragma unroll 1
                    for (int k = 0; k < 8; k++) {
                        const int b = 8 * (q + 4 * c) + k;
                        const float4 f = *reinterpret_cast<const float4 *>(scr + b * 128 + (((lane & 7) ^ (b & 7)) << 4));
                        const uint32_t sy = P.roff[b];
                        const int fl = P.flag[b];
                        const bool live = !(fl & 1);
                        const bool lastrow = (fl & 2) != 0;
                        int cnt = (f.x < 1e-10f) + (f.y < 1e-10f) + (f.z < 1e-10f) + (f.w < 1e-10f);
                        cnt = __reduce_add_sync(0xffffffffu, cnt);
                        if (cnt >= BW_N) {
                            if (live && lane == 0) atomicAdd(out.flags, 1.0);
                            cnt = BW_N - 1;
                        }
                        const float sub = (float)((double)cnt * 1e-10 / (double)(BW_N - cnt));
                        const float floorv = 1e-10f;
                        const float x0 = fmaxf(f.x - sub, floorv), x1 = fmaxf(f.y - sub, floorv),
                                    x2 = fmaxf(f.z - sub, floorv), x3 = fmaxf(f.w - sub, floorv);
                        const float inv = 1.0f / warp_sum((x0 + x1) + (x2 + x3));
                        if (live) {
                            float *row = tabr + 4 * lane;
                            red_add_v4_nc(row + sy, x0 * inv, x1 * inv, x2 * inv, x3 * inv);
                            red_add_v4_nc(row + (size_t)M * BW_N, x0 * inv, x1 * inv, x2 * inv, x3 * inv);
                            if (lastrow) red_add_v4_nc(row + (size_t)(M + 1) * BW_N, x0 * inv, x1 * inv, x2 * inv, x3 * inv);
                        }
                    }
                    red_add_f32_nc(tab0 + (size_t)(M + 1) * BW_N, glast);
                    named_bar_sync(1 + h, 256);     // the next tile's first step writes V into this buffer
                    if (Ti > 1 && h == 0) {
                        // tile done: Xi^T[j][i'] (lane = j = this thread's state) -> float64 counts xi[i'][j], coalesced over j;
                        // warp (q, c) takes the columns i' in [64c, 64c + 64)
                        mbar_wait(&bars->xi_done, ntile & 1);
                        tc_fence_after();
#pragma unroll 1
                        for (int c0 = 64 * c; c0 < 64 * c + 64; c0 += 32) {
                            uint32_t v[32];
                            tmem_ld_32x32(tmem + lane_base + C3_XI + c0, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int x = 0; x < 32; x++)
                                atomicAdd(out.xi + (c0 + x) * BW_N + i, (double)__uint_as_float(v[x]));
                        }
                        tc_fence_before();
                        mbar_arrive(&bars->xi_free);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->prep_empty[h][pb]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 17) tmem_dealloc(tmem, 512);
}

// Fold the fp32 table replicas into the float64 count buffer:
//   emis_num[k][i] += sum_r tab[r][k][i];  emis_den[i] += sum_k (that);  pi[i] += row M;
//   gam[i] += emis_den part - row M+1 (gamma_ref summed over t < T-1);  loglik, nseq.
static __global__ void bw_fold_kernel(const float *__restrict__ tab, int M, double *__restrict__ counts,
                                      const double *__restrict__ loglik, int64_t B) {
    const int i = threadIdx.x;   // 128 threads
    double *c_pi = counts;
    double *c_gam = c_pi + BW_N + BW_N * BW_N;
    double *c_num = c_gam + BW_N;
    double *c_den = c_num + (int64_t)M * BW_N;
    double *c_tail = c_den + BW_N;
    const size_t rstride = (size_t)(M + 2) * BW_N;
    double den = 0.0;
    for (int k = blockIdx.x; k < M + 2; k += gridDim.x) {
        double s = 0.0;
        for (int r = 0; r < BW_REPL; r++) s += (double)tab[r * rstride + (size_t)k * BW_N + i];
        if (k < M) {
            c_num[(int64_t)k * BW_N + i] += s;
            den += s;
        } else if (k == M) {
            atomicAdd(c_pi + i, s);
        } else {
            atomicAdd(c_gam + i, -s);
        }
    }
    atomicAdd(c_den + i, den);
    atomicAdd(c_gam + i, den);
    // log-likelihood sum
    double ll = 0.0;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + i; b < B; b += (int64_t)gridDim.x * blockDim.x) ll += loglik[b];
    ll = warp_sum(ll);
    if ((i & 31) == 0) atomicAdd(c_tail, ll);
    if (blockIdx.x == 0 && i == 0) atomicAdd(c_tail + 1, (double)B);
}

static long long *g_bw_trace = nullptr;   // development timeline (khmm_debug_bw_trace); NULL in production

// Optional kernel timing for bench.py's roofline (khmm_estep_tc_timing): a per-device ring of CUDA-event triples
// around the forward and backward launches of the last BW_RING calls.  Nothing is created or recorded unless armed.
constexpr int BW_RING = 64, BW_MAX_DEV = 64;
struct BwTiming {
    cudaEvent_t ev[BW_RING][3];
    bool created;
    unsigned long long calls;
};
static bool g_bw_timing = false;
static BwTiming g_bw_tm[BW_MAX_DEV];

struct BwWorkspace {
    float *stash, *scale, *tab, *Ar, *ATr;
    uint16_t *Ah, *Al, *ATh, *ATl;    // bf16 hi / lo halves of A and A^T (split-operand mode)
    double *loglik;
    size_t bytes;
};

// Which backward kernel: 3 = the transposed kernel with four quarter-tile pipelines (bw_backward_tc3_kernel, the default, both
// operand precisions); 1 = the round-1 thread-per-sequence kernel (bw_backward_tc_kernel, both precisions), kept as the
// independent cross-check of the default ($KHMM_BW_BACKWARD or khmm_debug_bw_backward_version select it).  Two further
// round-2 kernels were measured and removed: a transposed backward kernel with ONE tile per CTA (17.2-17.5 ms at config-4
// size against 17.0-17.7 ms for kernel 1) and a forward kernel on the quarter-tile structure (slower than the forward
// kernel above: DESIGN.md 3.5).
static int bw_backward_version() { return dev_switch(SW_BW_BACKWARD) == 1 ? 1 : 3; }

static BwWorkspace bw_layout(void *ws, int64_t B, int64_t T, int M) {
    auto up = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
    BwWorkspace w;
    size_t o = 0;
    unsigned char *p = (unsigned char *)ws;
    w.stash = (float *)(p + o);  o += up((size_t)((B + 127) / 128 * 128) * T * BW_N * 4);
    w.scale = (float *)(p + o);  o += up((size_t)((B + 127) / 128 * 128) * T * 4);
    w.tab = (float *)(p + o);    o += up((size_t)BW_REPL * (M + 2) * BW_N * 4);
    w.Ar = (float *)(p + o);     o += up((size_t)BW_N * BW_N * 4);
    w.ATr = (float *)(p + o);    o += up((size_t)BW_N * BW_N * 4);
    w.Ah = (uint16_t *)(p + o);  o += up((size_t)BW_N * BW_N * 2);
    w.Al = (uint16_t *)(p + o);  o += up((size_t)BW_N * BW_N * 2);
    w.ATh = (uint16_t *)(p + o); o += up((size_t)BW_N * BW_N * 2);
    w.ATl = (uint16_t *)(p + o); o += up((size_t)BW_N * BW_N * 2);
    w.loglik = (double *)(p + o); o += up((size_t)B * 8);
    w.bytes = o;
    return w;
}

static int make_tmap_stash(CUtensorMap *map, const float *stash, int64_t B, int64_t T, CUtensorMapSwizzle swz,
                           uint32_t box_rows = 128) {
    static PFN_encodeTiled encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -1;
        encode = (PFN_encodeTiled)fn;
    }
    // tile-major stash [ntiles * T][128 rows][128]
    cuuint64_t dims[3] = {(cuuint64_t)BW_N, 128, (cuuint64_t)((B + 127) / 128 * T)};
    cuuint64_t strides[2] = {(cuuint64_t)BW_N * 4, (cuuint64_t)128 * BW_N * 4};
    cuuint32_t box[3] = {32, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)stash, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace khmm

using namespace khmm;

extern "C" {

// development hook: device buffer of 2*8*32 int64 for the backward + forward timelines (NULL disarms it)
int khmm_debug_bw_trace(int64_t *dev_buf) {
    g_bw_trace = (long long *)dev_buf;
    return KHMM_OK;
}

int khmm_debug_bw_backward_version(int version) {
    KHMM_REQUIRE(version == 0 || version == 1 || version == 3, "debug_bw_backward_version: 0 (default), 1 or 3");
    return khmm_debug_switch("KHMM_BW_BACKWARD", version == 0 ? -1 : version);
}

int khmm_estep_tc_timing(int enable) {
    g_bw_timing = enable != 0;
    return KHMM_OK;
}

int khmm_estep_tc_kernel_ms(int calls_back, float *forward_ms, float *backward_ms) {
    int dev = 0;
    KHMM_CUDA_CHECK(cudaGetDevice(&dev));
    KHMM_REQUIRE(dev < BW_MAX_DEV && g_bw_tm[dev].created && g_bw_tm[dev].calls > 0,
                 "estep_tc_kernel_ms: no timed tensor-core E-step on this device (arm it with khmm_estep_tc_timing(1))");
    BwTiming &tm = g_bw_tm[dev];
    KHMM_REQUIRE(calls_back >= 0 && calls_back < BW_RING && (unsigned long long)calls_back < tm.calls,
                 "estep_tc_kernel_ms: only the last %d timed calls are kept", BW_RING);
    cudaEvent_t *ev = tm.ev[(tm.calls - 1 - calls_back) % BW_RING];
    KHMM_CUDA_CHECK(cudaEventSynchronize(ev[2]));
    if (forward_ms) KHMM_CUDA_CHECK(cudaEventElapsedTime(forward_ms, ev[0], ev[1]));
    if (backward_ms) KHMM_CUDA_CHECK(cudaEventElapsedTime(backward_ms, ev[1], ev[2]));
    return KHMM_OK;
}

int khmm_estep_tc_last_kernel_ms(float *forward_ms, float *backward_ms) {
    return khmm_estep_tc_kernel_ms(0, forward_ms, backward_ms);
}

size_t khmm_estep_tc_workspace_bytes(int64_t B, int64_t T, int N, int M) {
    if (N != BW_N || B < 0 || T < 1 || M < 1) return 0;
    return bw_layout(nullptr, B, T, M).bytes;
}

int khmm_estep_tc_discrete_f32(int64_t B, int64_t T, int N, int M, const float *pi, const float *A, const float *AT,
                               const float *Bsm, const void *obs, int obs_dtype, void *workspace,
                               size_t workspace_bytes, double *counts, float *alpha_out, float *scale_out,
                               khmm_stream_t stream) {
    return khmm_estep_tc_discrete_ragged_f32(B, T, N, M, pi, A, AT, Bsm, obs, obs_dtype, nullptr, workspace, workspace_bytes,
                                             counts, alpha_out, scale_out, stream);
}

int khmm_estep_tc_discrete_ragged_f32(int64_t B, int64_t T, int N, int M, const float *pi, const float *A, const float *AT,
                                      const float *Bsm, const void *obs, int obs_dtype, const int32_t *lengths,
                                      void *workspace, size_t workspace_bytes, double *counts, float *alpha_out,
                                      float *scale_out, khmm_stream_t stream) {
    return khmm_estep_tc_discrete_ex_f32(B, T, N, M, pi, A, AT, Bsm, obs, obs_dtype, lengths, workspace, workspace_bytes,
                                         counts, alpha_out, scale_out, KHMM_TC_TF32, stream);
}

int khmm_estep_tc_discrete_ex_f32(int64_t B, int64_t T, int N, int M, const float *pi, const float *A, const float *AT,
                                  const float *Bsm, const void *obs, int obs_dtype, const int32_t *lengths,
                                  void *workspace, size_t workspace_bytes, double *counts, float *alpha_out,
                                  float *scale_out, int operand_precision, khmm_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KHMM_REQUIRE(B >= 0 && T >= 1 && M >= 1, "estep_tc: bad shape");
    KHMM_REQUIRE(operand_precision == KHMM_TC_TF32 || operand_precision == KHMM_TC_BF16X3,
                 "estep_tc: operand_precision must be KHMM_TC_TF32 or KHMM_TC_BF16X3");
    const bool split = operand_precision != KHMM_TC_TF32;
    if (N != BW_N) {
        set_error("estep_tc: the fused tensor-core E-step covers N = %d (got %d)", BW_N, N);
        return KHMM_ERR_UNSUPPORTED;
    }
    if ((B + 127) / 128 * T > 0x7fffffff) {
        set_error("estep_tc: B and T must fit a 32-bit tensor-map coordinate");
        return KHMM_ERR_UNSUPPORTED;
    }
    if (B == 0) return KHMM_OK;
    KHMM_REQUIRE(pi && A && AT && Bsm && obs && workspace && counts, "estep_tc: NULL pointer argument");
    KHMM_REQUIRE(obs_dtype >= KHMM_U8 && obs_dtype <= KHMM_I64, "estep_tc: obs dtype code %d is not an integer type",
                 obs_dtype);
    KHMM_REQUIRE(((uintptr_t)workspace & 1023) == 0, "estep_tc: workspace must be 1024-byte aligned");
    BwWorkspace w = bw_layout(workspace, B, T, M);
    KHMM_REQUIRE(workspace_bytes >= w.bytes, "estep_tc: workspace too small");
    if (alpha_out) w.stash = alpha_out;   // caller-visible stash (tests): [ceil(B/128)][T][128][128], W_t = c_t alpha_t
    if (scale_out) w.scale = scale_out;
    KHMM_REQUIRE(((uintptr_t)w.stash & 15) == 0, "estep_tc: stash must be 16-byte aligned");

    CUtensorMap mapAT, mapATlo, mapA, mapAlo, mapStash, mapStashSt;
    int maperr;
    if (split) {
        bw_split_matrix_kernel<<<16, 256, 0, st>>>(A, w.Ah, w.Al, BW_N * BW_N);
        bw_split_matrix_kernel<<<16, 256, 0, st>>>(AT, w.ATh, w.ATl, BW_N * BW_N);
        maperr = make_tmap_bf16_rowmajor(&mapAT, w.ATh, BW_N, BW_N, 128) || make_tmap_bf16_rowmajor(&mapATlo, w.ATl, BW_N, BW_N, 128) ||
                 make_tmap_bf16_rowmajor(&mapA, w.Ah, BW_N, BW_N, 128) || make_tmap_bf16_rowmajor(&mapAlo, w.Al, BW_N, BW_N, 128);
    } else {
        bw_round_matrix_kernel<<<16, 256, 0, st>>>(A, w.Ar, BW_N * BW_N);
        bw_round_matrix_kernel<<<16, 256, 0, st>>>(AT, w.ATr, BW_N * BW_N);
        maperr = make_tmap_f32_rowmajor(&mapAT, w.ATr, BW_N, BW_N, 128) || make_tmap_f32_rowmajor(&mapA, w.Ar, BW_N, BW_N, 128);
        mapATlo = mapAT;
        mapAlo = mapA;
    }
    KHMM_CUDA_CHECK(cudaMemsetAsync(w.tab, 0, (size_t)BW_REPL * (M + 2) * BW_N * 4, st));
    if (maperr || make_tmap_stash(&mapStash, w.stash, B, T, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        make_tmap_stash(&mapStashSt, w.stash, B, T, CU_TENSOR_MAP_SWIZZLE_128B)) {
        set_error("estep_tc: tensor map creation failed");
        return KHMM_ERR_CUDA;
    }
    const int64_t ntiles = (B + 127) / 128;
    const int sms = sm_count();
    BwObs ob{obs, obs_dtype, M, T, lengths};
    cudaEvent_t *tev = nullptr;   // timing events of this call (armed by khmm_estep_tc_timing)
    {
        size_t smem = BW_TILE + 2 * (size_t)FW_EBUF + 6 * 128 * sizeof(float) + sizeof(FwdBars) + 1024;
        KHMM_CUDA_CHECK(cudaFuncSetAttribute(bw_forward_tc_kernel<KHMM_TC_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KHMM_CUDA_CHECK(cudaFuncSetAttribute(bw_forward_tc_kernel<KHMM_TC_BF16X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = (int)(ntiles < sms ? ntiles : sms);
        if (g_bw_timing) {
            int dev = 0;
            KHMM_CUDA_CHECK(cudaGetDevice(&dev));
            if (dev < BW_MAX_DEV) {
                BwTiming &tm = g_bw_tm[dev];
                if (!tm.created) {
                    for (int r = 0; r < BW_RING; r++)
                        for (int k = 0; k < 3; k++) KHMM_CUDA_CHECK(cudaEventCreate(&tm.ev[r][k]));
                    tm.created = true;
                }
                tev = tm.ev[tm.calls % BW_RING];
                tm.calls++;
            }
        }
        if (tev) KHMM_CUDA_CHECK(cudaEventRecord(tev[0], st));
        if (operand_precision == KHMM_TC_BF16X3)
            bw_forward_tc_kernel<KHMM_TC_BF16X3><<<grid, 384, smem, st>>>(mapAT, mapATlo, mapStashSt, pi, Bsm, ob, B, T, w.scale,
                                                                          w.loglik, g_bw_trace);
        else
            bw_forward_tc_kernel<KHMM_TC_TF32><<<grid, 384, smem, st>>>(mapAT, mapATlo, mapStashSt, pi, Bsm, ob, B, T, w.scale,
                                                                        w.loglik, g_bw_trace);
        KHMM_LAUNCH_CHECK("bw_forward_tc_kernel");
        if (tev) KHMM_CUDA_CHECK(cudaEventRecord(tev[1], st));
    }
    {
        size_t smem = 3 * (size_t)BW_TILE + 2 * (size_t)BW_ECHUNK + sizeof(BwdBars) + 1024;
        KHMM_CUDA_CHECK(cudaFuncSetAttribute(bw_backward_tc_kernel<KHMM_TC_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KHMM_CUDA_CHECK(cudaFuncSetAttribute(bw_backward_tc_kernel<KHMM_TC_BF16X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = (int)(ntiles < sms ? ntiles : sms);
        double *c_xi = counts + BW_N;
        double *c_tail = counts + BW_N + (size_t)BW_N * BW_N + BW_N + (size_t)M * BW_N + BW_N;
        BwdOut out{g_bw_trace, w.tab, c_xi, c_tail + 2};
        // kernel 3 forms table addresses with 32-bit adds: both tables must lie inside one 4 GB window each (else kernel 1)
        const bool windows_ok = within_4gb_window(Bsm, (size_t)M * BW_N * 4) &&
                                within_4gb_window(w.tab, (size_t)BW_REPL * (M + 2) * BW_N * 4);
        if (bw_backward_version() == 3 && windows_ok && M <= (1 << 22)) {
            CUtensorMap mapStashH;
            if (make_tmap_stash(&mapStashH, w.stash, B, T, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 64)) {
                set_error("estep_tc: tensor map creation failed");
                return KHMM_ERR_CUDA;
            }
            size_t smem3 = 6 * (size_t)BW_HTILE + 2 * BW3_PREP * sizeof(Bwd3Prep) + sizeof(Bwd3Bars) + 1024;
            {
                // the kernel re-splits its register file with setmaxnreg (16 compute warps x 104, 4 helper warps x 64): the
                // CTA's pool is what the launch allocates (registers per thread x 640), and a pool that is too small would
                // make setmaxnreg.inc wait forever -- refuse to launch instead
                cudaFuncAttributes fa;
                KHMM_CUDA_CHECK(split ? cudaFuncGetAttributes(&fa, bw_backward_tc3_kernel<KHMM_TC_BF16X3>)
                                      : cudaFuncGetAttributes(&fa, bw_backward_tc3_kernel<KHMM_TC_TF32>));
                if (fa.numRegs * BW3_THREADS < 512 * 104 + 128 * 64) {
                    set_error("estep_tc: bw_backward_tc3_kernel was compiled with %d registers per thread; its setmaxnreg split "
                              "needs 96", fa.numRegs);
                    return KHMM_ERR_UNSUPPORTED;
                }
            }
            if (split) {
                KHMM_CUDA_CHECK(cudaFuncSetAttribute(bw_backward_tc3_kernel<KHMM_TC_BF16X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
                bw_backward_tc3_kernel<KHMM_TC_BF16X3><<<grid, BW3_THREADS, smem3, st>>>(mapStashH, A, Bsm, ob, B, T, w.scale, out);
            } else {
                KHMM_CUDA_CHECK(cudaFuncSetAttribute(bw_backward_tc3_kernel<KHMM_TC_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
                bw_backward_tc3_kernel<KHMM_TC_TF32><<<grid, BW3_THREADS, smem3, st>>>(mapStashH, w.Ar, Bsm, ob, B, T, w.scale, out);
            }
            KHMM_LAUNCH_CHECK("bw_backward_tc3_kernel");
        } else if (operand_precision == KHMM_TC_BF16X3) {
            bw_backward_tc_kernel<KHMM_TC_BF16X3><<<grid, 384, smem, st>>>(mapA, mapAlo, mapStash, Bsm, ob, B, T, w.scale, out);
            KHMM_LAUNCH_CHECK("bw_backward_tc_kernel<bf16x3>");
        } else {
            bw_backward_tc_kernel<KHMM_TC_TF32><<<grid, 384, smem, st>>>(mapA, mapAlo, mapStash, Bsm, ob, B, T, w.scale, out);
            KHMM_LAUNCH_CHECK("bw_backward_tc_kernel");
        }
        if (tev) KHMM_CUDA_CHECK(cudaEventRecord(tev[2], st));
    }
    bw_fold_kernel<<<64, 128, 0, st>>>(w.tab, M, counts, w.loglik, B);
    KHMM_LAUNCH_CHECK("bw_fold_kernel");
    return KHMM_OK;
}

}  // extern "C"
