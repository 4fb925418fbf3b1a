// Tensor-core recursion for large state spaces (N a multiple of 128, N >= 128): the forward step
// alpha_{t-1} . A (abstract_hmm.py:266-270) and the backward step A . (b . beta_{t+1}) (:339-349) of
// a whole batch are one (B x N) x (N x N) contraction per time step.  Each step is ONE launch of a
// tcgen05 kernel (kind::tf32, fp32 accumulators in TMEM, operands staged by TMA with the 128-byte
// swizzle) whose epilogue fuses the emission evaluation, the scaling and the row sums:
//
//     W_t = (W_{t-1} M) . b_t . r_{t-1},      r_{t-1} = 1 / sum_j W_{t-1}(j),
//
// M = A with reversed time gives the (self-scaled) backward recursion in the form w_t = b_t . beta_t,
// exactly as in smalln.cu.  W_t is the reference's *unnormalised* alpha_t (its row sum is the scale
// factor c_t), so normalisation needs no second pass over a row that is split across CTAs: every
// column tile writes its partial row sum and the next step's epilogue adds them up.  Forward and
// backward share a launch (blockIdx.z), log P accumulates in float64 from the c_t.
//
// Warp roles per CTA (192 threads): warps 0-3 epilogue (thread = TMEM lane = sequence), warp 4 TMA
// producer, warp 5 TMEM allocator + single-thread MMA issuer.  4-stage smem ring of
// (128 x 32 fp32) W tiles and (128 x 32 fp32) matrix tiles.
//
// Precision: tf32 multiplies (10-bit mantissa, inputs truncated by the tensor core), fp32
// accumulation.  Measured against the float64 oracle at N=256..1024 the log-likelihood agrees to
// ~1e-6 relative (tests/test_gpu_parity.py states the bound); alpha itself is only good to ~1e-3,
// which is why the float32 SIMT kernels remain the parity path for alpha/beta outputs.
#include "common.cuh"
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
};

// Round to nearest tf32 (10-bit mantissa).  The tensor core TRUNCATES fp32 operands to tf32, which
// biases every product low (measured: log-likelihood off by -6.6e-4 per step); operands that are
// already tf32-representable pass through exactly, so rounding them here makes the error unbiased.
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
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

// step 0: W_0 = start . b(o_first): start = pi (forward, t = 0) or ones (backward, t = T-1)
template <bool GAUSS>
__global__ void tc_init_kernel(TcDir df, TcDir db, TcEmis em, const float *__restrict__ pi, int64_t B, int64_t T, int N) {
    const int dir = blockIdx.y;
    const TcDir d = dir ? db : df;
    const int64_t b = blockIdx.x;
    const int NT = N / TC_BN;
    const int64_t t = dir ? T - 1 : 0;
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
        float v = (dir ? 1.f : pi[j]) * e;
        d.W[(b * 2 + 0) * N + j] = round_tf32(v);
        float s = warp_sum(v);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) d.psum[(b * 2 + 0) * NT + nt] = red[0] + red[1] + red[2] + red[3];
    }
    if (threadIdx.x == 0) d.logacc[b] = 0.0;
}

// closing: forward log P = logacc + log c_{T-1}; backward log P = logacc + log(pi . w_0)
__global__ void tc_final_kernel(TcDir df, TcDir db, const float *__restrict__ pi, int64_t B, int64_t T, int N,
                                double *__restrict__ loglik, double *__restrict__ loglik_bwd) {
    const int dir = blockIdx.y;
    const TcDir d = dir ? db : df;
    const int64_t b = blockIdx.x;
    const int slot = (int)((T - 1) & 1);
    const float *w = d.W + (b * 2 + slot) * N;
    float s = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) s += dir ? w[j] * pi[j] : w[j];
    __shared__ float red[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) tot += red[k];
        double ll = d.logacc[b] + log((double)tot);
        if (dir) { if (loglik_bwd) loglik_bwd[b] = ll; } else loglik[b] = ll;
    }
}

static size_t tc_workspace_floats(int64_t B, int N) {
    int NT = N / TC_BN;
    int64_t Bp = (B + TC_BM - 1) / TC_BM * TC_BM;
    // per direction: W [Bp][2][N], psum [Bp][2][NT], logacc [Bp] doubles (2 floats each)
    return (size_t)2 * ((size_t)Bp * 2 * N + (size_t)Bp * 2 * NT + (size_t)Bp * 2) + (size_t)2 * N * N + 64;
}

template <bool GAUSS>
static int run_tc(int64_t B, int64_t T, int N, const float *pi, const float *A, const float *AT, TcEmis em, void *ws,
                  size_t ws_bytes, double *loglik, double *loglik_bwd, cudaStream_t st) {
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
        d[k].psum = p;   p += (size_t)Bp * 2 * NT;
        d[k].logacc = (double *)p; p += (size_t)Bp * 2;
    }
    float *Ar = p;  p += (size_t)N * N;     // tf32-rounded copies of A and AT
    float *ATr = p; p += (size_t)N * N;
    tc_round_matrix_kernel<<<64, 256, 0, st>>>(A, Ar, N * N);
    tc_round_matrix_kernel<<<64, 256, 0, st>>>(AT, ATr, N * N);
    const int ndir = loglik_bwd ? 2 : 1;
    tc_init_kernel<GAUSS><<<dim3((unsigned)B, ndir), 128, 0, st>>>(d[0], d[1], em, pi, B, T, N);
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
            cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)Bp};
            cuuint64_t strides[1] = {(cuuint64_t)2 * N * sizeof(float)};
            cuuint32_t box[2] = {TC_BK, TC_BM};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = encode(&mapX[k][slot], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)(d[k].W + (size_t)slot * N),
                                dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
                return KHMM_ERR_CUDA;
            }
        }
    // 256-wide column tiles when N allows: 25 % less L2 operand traffic per MAC and half the CTAs
    // (small batches -- e.g. a strong-scaling shard -- keep 128-wide tiles so that more SMs get a CTA)
    const int BN = (N % 256 == 0 && (Bp / TC_BM) * (N / 256) * ndir >= sm_count()) ? 256 : 128;
    // forward: out[j] = sum_i W[i] A[i][j]  -> B-operand rows j, K index i: AT[j][i]
    // backward: out[i] = sum_j A[i][j] w[j] -> B-operand rows i, K index j: A[i][j]
    if (make_tmap_f32_rowmajor(&mapM[0], ATr, N, N, BN) || make_tmap_f32_rowmajor(&mapM[1], Ar, N, N, BN)) {
        set_error("tensor map creation for the transition matrix failed");
        return KHMM_ERR_CUDA;
    }
    size_t smem = TC_STAGES * (size_t)(TC_TILE_BYTES + BN * TC_BK * 4) + sizeof(TcBars) + 3 * BN * sizeof(float) + 1024;
    auto kern = BN == 256 ? tc_step_kernel<GAUSS, 256> : tc_step_kernel<GAUSS, 128>;
    KHMM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(Bp / TC_BM), N / BN, ndir);
    for (int step = 1; step < T; step++) {
        int src = (step - 1) & 1;
        kern<<<grid, 192, smem, st>>>(mapX[0][src], mapX[1][src], mapM[0], mapM[1], d[0], d[1], em, B, T, N, step);
    }
    KHMM_LAUNCH_CHECK("tc_step_kernel");
    tc_final_kernel<<<dim3((unsigned)B, ndir), 256, 0, st>>>(d[0], d[1], pi, B, T, N, loglik, loglik_bwd);
    KHMM_LAUNCH_CHECK("tc_final_kernel");
    return KHMM_OK;
}

}  // namespace khmm

using namespace khmm;

extern "C" {

size_t khmm_fwdbwd_loglik_tc_workspace_bytes(int64_t B, int N) {
    if (N < TC_BN || N % TC_BN != 0) return 0;
    return tc_workspace_floats(B, N) * sizeof(float);
}

int khmm_fwdbwd_loglik_tc_gauss_f32(int64_t B, int64_t T, int N, const float *pi, const float *A, const float *AT,
                                    const float *mean, const float *sd, const void *obs, int obs_dtype, void *workspace,
                                    size_t workspace_bytes, double *loglik, double *loglik_bwd, khmm_stream_t stream) {
    KHMM_REQUIRE(B == 0 || (mean && sd && obs), "fwdbwd_loglik_tc_gauss: NULL pointer argument");
    KHMM_REQUIRE(obs_dtype == KHMM_F32 || obs_dtype == KHMM_F64, "gauss: obs dtype code %d is not f32/f64", obs_dtype);
    TcEmis em{1, mean, sd, obs, obs_dtype, 0};
    return run_tc<true>(B, T, N, pi, A, AT, em, workspace, workspace_bytes, loglik, loglik_bwd, (cudaStream_t)stream);
}

int khmm_fwdbwd_loglik_tc_discrete_f32(int64_t B, int64_t T, int N, int M, const float *pi, const float *A,
                                       const float *AT, const float *Bsm, const void *obs, int obs_dtype,
                                       void *workspace, size_t workspace_bytes, double *loglik, double *loglik_bwd,
                                       khmm_stream_t stream) {
    KHMM_REQUIRE(B == 0 || (Bsm && obs && M >= 1), "fwdbwd_loglik_tc_discrete: NULL pointer argument");
    KHMM_REQUIRE(obs_dtype >= KHMM_U8 && obs_dtype <= KHMM_I64, "discrete: obs dtype code %d is not an integer type",
                 obs_dtype);
    TcEmis em{0, Bsm, nullptr, obs, obs_dtype, M};
    return run_tc<false>(B, T, N, pi, A, AT, em, workspace, workspace_bytes, loglik, loglik_bwd, (cudaStream_t)stream);
}

}  // extern "C"
