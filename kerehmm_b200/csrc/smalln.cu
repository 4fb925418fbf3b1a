// Fused forward-backward log-likelihood for small state spaces (N <= 16), fp32.
//
// One THREAD per (sequence, direction): the state vector (alpha-hat / b.beta-hat) and, for
// N <= 8, the whole transition matrix live in registers, the emission evaluation (Gaussian
// pdf or discrete gather) is fused into the step, nothing but the observations is read from HBM
// and nothing but 8 bytes per sequence and direction is written.  Even warps run the forward
// recursion (abstract_hmm.py:250-285), odd warps the backward recursion (:327-351) of the same 32
// sequences, concurrently -- the backward pass is self-scaled, so it does not wait for the
// forward scale factors:
//     forward : u_0 = pi.b_0/c_0,        u_t = ((u_{t-1} A) . b_t)/c_t,   log P = sum log c_t
//     backward: w_{L-1} = b_{L-1}/s,     w_t = ((A w_{t+1}) . b_t)/s_t,   log P = sum log s_t + log(pi.w_0)
// (w_t = b_t . beta_t up to scaling).  Both are "row vector times matrix, times emission,
// normalise" -- with A^T and reversed time for the backward pass -- so one code path serves both.
//
// (N = 5..8 takes fwdbwd_mma8_kernel further down, which runs the matrix-vector products on the tensor cores; the
// thread-per-sequence kernel described here serves the other N and stays selectable as its cross-check.)
//
// Blackwell specifics: all FP32 work is issued as packed FFMA2/FMUL2/FADD2 (fma.rn.f32x2, two
// states per instruction) -- the kernel is bound by FP32 issue, not by HBM; the
// matrix-vector product keeps one accumulation chain per pair of OUTPUT states (broadcast input
// state times a packed pair of matrix entries); exp2 is a single MUFU instruction (ex2.approx);
// observations are prefetched one 8-step block ahead so their latency never sits on the
// recursion's dependency chain.
//
// Scaling (round 2): exact powers of two, taken from the exponent of the row sum every OTHER step and applied one step
// late, off the recursion's dependency chain; log P = ln 2 * (sum of the exponents, an integer) + log(final sum).  No
// reciprocal, no mantissa bookkeeping, and half the scale / row-sum arithmetic of scaling at every step.
#include "common.cuh"

namespace khmm {

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int NP, bool GAUSS, bool MX_SMEM>
struct SmallStep {
    static constexpr int H = NP / 2;
    // mx2[y][k] = (M[y][2k], M[y][2k+1]) with M = A (forward) or A^T (backward)
    float2 mx2[MX_SMEM ? 1 : NP][MX_SMEM ? 1 : H];
    const float2 *mxs;     // shared-memory copy when MX_SMEM
    float2 nm[H], kk[H], lc[H];  // gaussian: -mean, -0.5*log2(e)/sd^2, -log2(sd*sqrt(2pi)) per state pair
    const float *Bsm;
    int N;

    __device__ __forceinline__ float2 mat(int y, int k) const { return MX_SMEM ? mxs[y * H + k] : mx2[y][k]; }

    // emission values of the NP states for one observation
    __device__ __forceinline__ void emis(float x, int sym, float2 (&bv)[H]) const {
        if (GAUSS) {
            const float2 xx = make_float2(x, x);
#pragma unroll
            for (int k = 0; k < H; k++) {
                float2 d = __fadd2_rn(xx, nm[k]);
                float2 a = __ffma2_rn(__fmul2_rn(d, d), kk[k], lc[k]);
                bv[k] = make_float2(ex2_approx(a.x), ex2_approx(a.y));
            }
        } else {
            const float *row = Bsm + (int64_t)sym * N;
#pragma unroll
            for (int k = 0; k < H; k++)
                bv[k] = make_float2(2 * k < N ? __ldg(row + 2 * k) : 0.f, 2 * k + 1 < N ? __ldg(row + 2 * k + 1) : 0.f);
        }
    }

    // w M: input states in ONE accumulation chain per output pair (the round-1 version split even and odd inputs into two
    // chains and paid a packed add per output pair to merge them), written input-state-major so that the H chains
    // advance together and consecutive instructions are independent (the compiler kept output-pair-major source order
    // and interleaved only two chains: 0.355 against 0.340 ms at config-2 size)
    __device__ __forceinline__ void matvec(const float2 (&w)[H], float2 (&out)[H]) const {
        if constexpr (MX_SMEM) {        // N = 9..16: eight chains at once would only add to the register pressure
#pragma unroll
            for (int k = 0; k < H; k++) {
                float2 c = __fmul2_rn(make_float2(w[0].x, w[0].x), mat(0, k));
                c = __ffma2_rn(make_float2(w[0].y, w[0].y), mat(1, k), c);
#pragma unroll
                for (int y2 = 1; y2 < H; y2++) {
                    c = __ffma2_rn(make_float2(w[y2].x, w[y2].x), mat(2 * y2, k), c);
                    c = __ffma2_rn(make_float2(w[y2].y, w[y2].y), mat(2 * y2 + 1, k), c);
                }
                out[k] = c;
            }
        } else {
#pragma unroll
            for (int k = 0; k < H; k++) out[k] = __fmul2_rn(make_float2(w[0].x, w[0].x), mat(0, k));
#pragma unroll
            for (int k = 0; k < H; k++) out[k] = __ffma2_rn(make_float2(w[0].y, w[0].y), mat(1, k), out[k]);
#pragma unroll
            for (int y2 = 1; y2 < H; y2++) {
#pragma unroll
                for (int k = 0; k < H; k++) out[k] = __ffma2_rn(make_float2(w[y2].x, w[y2].x), mat(2 * y2, k), out[k]);
#pragma unroll
                for (int k = 0; k < H; k++)
                    out[k] = __ffma2_rn(make_float2(w[y2].y, w[y2].y), mat(2 * y2 + 1, k), out[k]);
            }
        }
    }

    // Steps come in pairs, and the scale factor is an exact POWER OF TWO applied one step late:
    //     step A:  w = (w M) . b                       s = sum(w)  (off the dependency chain);  e = exponent(s)
    //     step B:  w = (w M) . (b 2^-e)                log2 P accumulates e as an integer
    // The recursion is linear, so any positive scale sequence gives the same likelihood; powers of two make the
    // bookkeeping exact and free of MUFU work, and scaling every other step halves the scale and row-sum arithmetic
    // (52 packed FMA-pipe instructions per step at N = 8 against 64 in round 1).  Between two rescalings the vector
    // shrinks by at most three emission factors, far inside the fp32 range for any emission above 1e-12.
    __device__ __forceinline__ void step_plain(float2 (&w)[H], const float2 (&bv)[H]) const {
        float2 out[H];
        matvec(w, out);
#pragma unroll
        for (int k = 0; k < H; k++) w[k] = __fmul2_rn(out[k], bv[k]);
    }
    __device__ __forceinline__ void step_scaled(float2 (&w)[H], const float2 (&bv)[H], float sc) const {
        const float2 s2 = make_float2(sc, sc);
        float2 out[H];
        matvec(w, out);
#pragma unroll
        for (int k = 0; k < H; k++) w[k] = __fmul2_rn(out[k], __fmul2_rn(bv[k], s2));
    }
    __device__ __forceinline__ float hsum(const float2 (&val)[H]) const {
        float2 s2 = val[0];
        if constexpr (H >= 2) s2 = __fadd2_rn(s2, val[1]);
        if constexpr (H >= 4) s2 = __fadd2_rn(s2, __fadd2_rn(val[2], val[3]));
        if constexpr (H >= 8) s2 = __fadd2_rn(s2, __fadd2_rn(__fadd2_rn(val[4], val[5]), __fadd2_rn(val[6], val[7])));
        return s2.x + s2.y;
    }
    // 2^-e with e = the exponent of s (s = 0 or subnormal: 2^127, the vector is zero anyway); esum += e
    __device__ __forceinline__ float pow2_rescale(float s, int &esum) const {
        const int e = (int)((__float_as_uint(s) >> 23) & 0xffu) - 127;
        esum += e;
        return __uint_as_float((uint32_t)(127 - e) << 23);
    }

};

constexpr int kBlk = 8;  // steps per observation prefetch block

template <int NP, bool GAUSS, bool MX_SMEM, int Q>
__global__ void __launch_bounds__(64)
fwdbwd_small_kernel(int64_t B, int64_t T, int N, int M, const float *__restrict__ pi, const float *__restrict__ A,
                    const float *__restrict__ p0, const float *__restrict__ p1, const void *__restrict__ obs,
                    int obs_dtype, const int32_t *__restrict__ lengths, double *__restrict__ loglik,
                    double *__restrict__ loglik_bwd) {
    // Q sequences per thread: their recursions are independent, so the compiler interleaves the two
    // dependency chains and one warp keeps the FMA pipe busy on its own (the model constants are
    // shared between them).
    constexpr int H = NP / 2;
    __shared__ float2 mx_sh[MX_SMEM ? 2 * H * NP : 1];  // [dir][y][k]
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const bool two = loglik_bwd != nullptr;
    const int dir = two ? (int)(warp & 1) : 0;  // warp-uniform: 0 forward, 1 backward
    const int64_t bfirst = (two ? (warp >> 1) : warp) * (32 * Q) + lane;   // sequences bfirst + 32*q

    auto m_elem = [&](int d, int y, int x) -> float {  // M[y][x], zero padded
        if (y >= N || x >= N) return 0.f;
        return d ? __ldg(A + x * N + y) : __ldg(A + y * N + x);
    };
    SmallStep<NP, GAUSS, MX_SMEM> S;
    S.N = N;
    S.Bsm = p0;
    S.mxs = nullptr;
    if (MX_SMEM) {
        for (int k = threadIdx.x; k < 2 * H * NP; k += blockDim.x) {
            int d = k / (H * NP), y = (k / H) % NP, kk2 = k % H;
            mx_sh[k] = make_float2(m_elem(d, y, 2 * kk2), m_elem(d, y, 2 * kk2 + 1));
        }
        __syncthreads();
        S.mxs = mx_sh + dir * H * NP;
    } else {
#pragma unroll
        for (int y = 0; y < NP; y++)
#pragma unroll
            for (int k = 0; k < H; k++) S.mx2[y][k] = make_float2(m_elem(dir, y, 2 * k), m_elem(dir, y, 2 * k + 1));
    }
    float2 pv[H];
#pragma unroll
    for (int k = 0; k < H; k++) {
        pv[k] = make_float2(2 * k < N ? __ldg(pi + 2 * k) : 0.f, 2 * k + 1 < N ? __ldg(pi + 2 * k + 1) : 0.f);
        if (GAUSS) {
            float m[2], kq[2], l[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                int j = 2 * k + q;
                float sd = j < N ? __ldg(p1 + j) : 1.f;
                m[q] = j < N ? -__ldg(p0 + j) : 0.f;
                kq[q] = j < N ? -0.5f * 1.4426950408889634f / (sd * sd) : 0.f;
                l[q] = j < N ? -log2f(sd * 2.5066282746310002f) : -INFINITY;  // padded states emit 0
            }
            S.nm[k] = make_float2(m[0], m[1]);
            S.kk[k] = make_float2(kq[0], kq[1]);
            S.lc[k] = make_float2(l[0], l[1]);
        }
    }
    if (bfirst >= B) return;
    int len[Q];
    int64_t base[Q];
    int lmax = 0;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        int64_t b = bfirst + 32 * q;
        len[q] = b < B ? (lengths ? lengths[b] : (int)T) : 0;
        base[q] = (b < B ? b : bfirst) * T;
        lmax = max(lmax, len[q]);
    }

    // observation of recursion step `s` of sequence q (time runs backwards for the backward pass)
    auto load_x = [&](int q, int s) -> float {
        int t = dir ? len[q] - 1 - s : s;
        if (GAUSS) return load_real<float>(obs, obs_dtype, base[q] + t);
        int sym = load_symbol(obs, obs_dtype, base[q] + t);
        sym = sym < 0 ? 0 : (sym >= M ? M - 1 : sym);
        return __int_as_float(sym);
    };

    float2 u[Q][H];
    float sc[Q];        // pending power-of-two scale factor (applied by the next scaled step)
    int esum[Q];        // log2 of the product of the scale factors taken out so far, negated
#pragma unroll
    for (int q = 0; q < Q; q++) {   // step 0: start vector (pi forward, ones backward) times the first emission
        float2 bv[H];
        float x0 = len[q] > 0 ? load_x(q, 0) : 0.f;
        S.emis(x0, __float_as_int(x0), bv);
#pragma unroll
        for (int k = 0; k < H; k++) {
            float2 st = dir ? make_float2(2 * k < N ? 1.f : 0.f, 2 * k + 1 < N ? 1.f : 0.f) : pv[k];
            u[q][k] = __fmul2_rn(st, bv[k]);
        }
        esum[q] = 0;
        sc[q] = S.pow2_rescale(S.hsum(u[q]), esum[q]);
    }
    float xs[Q][kBlk], xn[Q][kBlk];
#pragma unroll
    for (int q = 0; q < Q; q++)
#pragma unroll
        for (int i = 0; i < kBlk; i++) xs[q][i] = (1 + i < len[q]) ? load_x(q, 1 + i) : 0.f;
    for (int s0 = 1; s0 < lmax; s0 += kBlk) {
#pragma unroll
        for (int q = 0; q < Q; q++)
#pragma unroll
            for (int i = 0; i < kBlk; i++) xn[q][i] = (s0 + kBlk + i < len[q]) ? load_x(q, s0 + kBlk + i) : 0.f;
        bool full = true;
#pragma unroll
        for (int q = 0; q < Q; q++) full = full && (s0 + kBlk <= len[q]);
        if (full) {
#pragma unroll
            for (int i = 0; i < kBlk; i += 2) {
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    float2 bv[H];
                    // scaled step (takes the pending factor), then a plain step whose row sum gives the next factor
                    S.emis(xs[q][i], __float_as_int(xs[q][i]), bv);
                    S.step_scaled(u[q], bv, sc[q]);
                    S.emis(xs[q][i + 1], __float_as_int(xs[q][i + 1]), bv);
                    S.step_plain(u[q], bv);
                    sc[q] = S.pow2_rescale(S.hsum(u[q]), esum[q]);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < kBlk; i++) {
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    if (s0 + i < len[q]) {      // the ragged tail rescales at every step
                        float2 bv[H];
                        S.emis(xs[q][i], __float_as_int(xs[q][i]), bv);
                        S.step_scaled(u[q], bv, sc[q]);
                        sc[q] = S.pow2_rescale(S.hsum(u[q]), esum[q]);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < Q; q++) {
#pragma unroll
            for (int i = 0; i < kBlk; i++) xs[q][i] = xn[q][i];
        }
    }
#pragma unroll
    for (int q = 0; q < Q; q++) {
        int64_t b = bfirst + 32 * q;
        if (b >= B) continue;
        // closing dot product: forward sum(u), backward pi . w_0 -- of the vector as it stands (the pending factor was
        // counted in esum but not applied yet, so it is applied here)
        float2 f2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < H; k++) {
            float2 w = dir ? pv[k] : make_float2(2 * k < N ? 1.f : 0.f, 2 * k + 1 < N ? 1.f : 0.f);
            f2 = __ffma2_rn(u[q][k], w, f2);
        }
        const float fin = (f2.x + f2.y) * sc[q];
        double ll = (double)esum[q] * 0.6931471805599453 + log((double)fin);
        if (!(fin > 0.f) || !(fin < INFINITY)) ll = __longlong_as_double(0x7ff8000000000000ll);  // underflow -> NaN (Q15)
        (dir ? loglik_bwd : loglik)[b] = ll;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// N in 5..8 (round 2): the matrix-vector products of 16 sequences as ONE warp-level tensor-core tile.
//
// The kernel above is bound by the FP32 pipe, and only 62 % of what it issues there is the matrix-vector product (the rest
// is the emission density, the scaling and the row sum).  Here a warp owns 16 sequences of one direction and the product
// (16 x 8) . (8 x 8) is one mma.sync.m16n8k8 (tf32 operands, fp32 accumulate) on split operands
//     w M  ~  w_hi M_hi + (w_lo M_hi + w_hi M_lo)      (w_hi = the upper 19 bits of w, w_lo = w - w_hi exactly)
// with the two correction terms -- 2^-11 of the product, so bf16 operands leave 2^-19 of it -- as ONE bf16 m16n8k16
// whose K index runs over (w_lo, w_hi) against (M_hi; M_lo): two tensor-core instructions per step and fp32-class
// accuracy (3e-8 relative on log P against float64 at config-2 size).  The fragment layouts make
// the recursion shuffle-free: lane (g = lane / 4, c = lane % 4) receives D[g][2c], D[g][2c+1], D[g+8][2c], D[g+8][2c+1]
// and has to supply A[g][c], A[g+8][c], A[g][c+4], A[g+8][c+4] -- so the K index is permuted (k = c -> state 2c,
// k = c + 4 -> state 2c + 1, the rows of the matrix operand permuted to match) and a lane's outputs ARE its next inputs.
// A lane therefore owns the state pair (2c, 2c+1) of sequences g and g + 8: two packed emission evaluations per step, two
// packed multiplies, and every other step a third mma against a matrix of ones that returns the row sums (only their
// exponent is used: the scale factors stay exact powers of two, as above).  What is left on the FP32 pipe is a quarter of
// the kernel above per sequence; the step costs the SUM of what its instruction classes cost the scheduler
// (profiles/r2_c2_mma8_ablation.txt): an mma.sync holds it for ~8 cycles, a packed FP32 instruction ~2.2.
__device__ __forceinline__ void mma_m16n8k8_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                 uint32_t b0, uint32_t b1, const float (&c)[4]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
__device__ __forceinline__ void mma_m16n8k16_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                  uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// two floats -> packed bf16 pair, `lo` in the lower half (the lower k index of a fragment register)
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
#ifndef KHMM_MMA8_BF16CORR
#define KHMM_MMA8_BF16CORR 1     /* 0: the two correction terms as two more tf32 mma (0.295 against 0.288 ms at config 2) */
#endif
__device__ __forceinline__ uint32_t tf32_rna(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// one observation as the float the emission evaluation takes: the value (Gaussian) or the bits of the clamped symbol
template <int DT>
__device__ __forceinline__ float obs_elem(const void *obs, int64_t idx, int M) {
    if (DT == KHMM_F32) return __ldg((const float *)obs + idx);
    if (DT == KHMM_F64) return (float)__ldg((const double *)obs + idx);
    int sym;
    if (DT == KHMM_U8) sym = (int)__ldg((const uint8_t *)obs + idx);
    else if (DT == KHMM_U16) sym = (int)__ldg((const uint16_t *)obs + idx);
    else if (DT == KHMM_I32) sym = __ldg((const int32_t *)obs + idx);
    else sym = (int)__ldg((const long long *)obs + idx);
    return __int_as_float(min(max(sym, 0), M - 1));
}

// DIR (0 forward, 1 backward) and the observation type are compile-time: the observation loads of the unrolled path are
// then plain loads at immediate offsets from one running pointer per sequence (the first version computed a 64-bit
// address, a type dispatch and a length predicate per element: half of the instructions it issued).
#ifndef KHMM_MMA8_BLK
#define KHMM_MMA8_BLK 8
#endif
constexpr int kBlkM = KHMM_MMA8_BLK;  // steps per observation prefetch block of the warp-tile kernel

// VEC: every sequence has the full length T, T is a multiple of 8 and the rows are 16-byte aligned (the host checks) --
// the observations of a block are then two 16-byte loads per sequence, a quarter of the L1 data-pipe wavefronts of eight
// 4-byte loads (one wavefront per 128-byte line and instruction; that pipe was 77 % busy and set the pace), the blocks
// are aligned to multiples of 8 steps and there is no ragged tail.
template <bool GAUSS, int DT, int DIR, bool VEC>
__device__ __forceinline__ void mma8_body(int64_t B, int64_t T, int N, int M, const float *__restrict__ pi,
                                          const float *__restrict__ A, const float *__restrict__ p0,
                                          const float *__restrict__ p1, const void *__restrict__ obs,
                                          const int32_t *__restrict__ lengths, double *__restrict__ out, int64_t tile) {
    const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const int j0 = 2 * c, j1 = 2 * c + 1;                   // the state pair of this lane

    // matrix operand: B[k][n] with n = g the output state; k = c -> input state 2c, k = c + 4 -> input state 2c + 1
    auto m_elem = [&](int y, int x) -> float {              // M[y][x], M = A forward, A^T backward, zero padded
        if (y >= N || x >= N) return 0.f;
        return DIR ? __ldg(A + x * N + y) : __ldg(A + y * N + x);
    };
    const float m0 = m_elem(j0, g), m1 = m_elem(j1, g);
    const uint32_t bh0 = tf32_rna(m0), bh1 = tf32_rna(m1);
    const uint32_t bl0 = tf32_rna(m0 - __uint_as_float(bh0)), bl1 = tf32_rna(m1 - __uint_as_float(bh1));
    // the two correction terms w_lo M_hi + w_hi M_lo as ONE bf16 m16n8k16 (k < 8: w_lo against M_hi, k >= 8: w_hi against
    // M_lo; natural state order): they are 2^-11 of the product, so bf16's 2^-9 leaves 2^-19 of it
    const uint32_t bc0 = pack_bf16(__uint_as_float(bh0), __uint_as_float(bh1));
    const uint32_t bc1 = pack_bf16(m0 - __uint_as_float(bh0), m1 - __uint_as_float(bh1));
    const uint32_t kOne = 0x3f800000u;
    const float zero4[4] = {0.f, 0.f, 0.f, 0.f};

    const float2 pv = make_float2(j0 < N ? __ldg(pi + j0) : 0.f, j1 < N ? __ldg(pi + j1) : 0.f);
    const float2 live = make_float2(j0 < N ? 1.f : 0.f, j1 < N ? 1.f : 0.f);
    float2 nm = make_float2(0.f, 0.f), kk = nm, lc = nm;
    if (GAUSS) {
        float mm[2], kq[2], l[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int j = 2 * c + q;
            const float sd = j < N ? __ldg(p1 + j) : 1.f;
            mm[q] = j < N ? -__ldg(p0 + j) : 0.f;
            kq[q] = j < N ? -0.5f * 1.4426950408889634f / (sd * sd) : 0.f;
            l[q] = j < N ? -log2f(sd * 2.5066282746310002f) : -INFINITY;   // padded states emit 0
        }
        nm = make_float2(mm[0], mm[1]);
        kk = make_float2(kq[0], kq[1]);
        lc = make_float2(l[0], l[1]);
    }
    auto emis = [&](float x) -> float2 {                    // the two emission values of this lane for one observation
        if (GAUSS) {
            const float2 d = __fadd2_rn(make_float2(x, x), nm);
            const float2 a = __ffma2_rn(__fmul2_rn(d, d), kk, lc);
            return make_float2(ex2_approx(a.x), ex2_approx(a.y));
        }
        const float *row = p0 + (int64_t)__float_as_int(x) * N;
        return make_float2(j0 < N ? __ldg(row + j0) : 0.f, j1 < N ? __ldg(row + j1) : 0.f);
    };

    int len[2];
    int64_t first[2];                                       // element index of recursion step 0
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int64_t b = tile * 16 + g + 8 * q;
        len[q] = b < B ? (lengths ? lengths[b] : (int)T) : 0;
        first[q] = (b < B ? b : tile * 16) * T + (DIR ? len[q] - 1 : 0);
    }
    int lmax = max(len[0], len[1]);
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    // observation of recursion step s (time runs backwards for the backward pass); 0 past the end of the sequence
    auto load_x = [&](int q, int s) -> float {
        return s < len[q] ? obs_elem<DT>(obs, first[q] + (DIR ? -s : s), M) : 0.f;
    };

    // the vector in fragment order: w[0] = w(g, 2c), w[1] = w(g+8, 2c), w[2] = w(g, 2c+1), w[3] = w(g+8, 2c+1)
    // (scalars, not packed pairs: the accumulator pairs states of one sequence, the A fragment pairs sequences of one
    // state, and packed results would have to be moved between register pairs at every step)
    float w[4];
    float sc[2];
    int esum[2] = {0, 0};
    uint32_t h[4], l[4];                                     // operands: upper 19 bits and the exact remainder
    auto split = [&]() {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            h[i] = __float_as_uint(w[i]) & 0xffffe000u;
            l[i] = __float_as_uint(w[i] - __uint_as_float(h[i]));
        }
    };
    auto product = [&](float (&d)[4]) {                      // d = (w M) of both sequences, from h / l
        mma_m16n8k8_tf32(d, h[0], h[1], h[2], h[3], bh0, bh1, zero4);
#if KHMM_MMA8_BF16CORR
        // fragment order of w / h / l: [0] = (g, 2c), [1] = (g+8, 2c), [2] = (g, 2c+1), [3] = (g+8, 2c+1)
        mma_m16n8k16_bf16(d, pack_bf16(__uint_as_float(l[0]), __uint_as_float(l[2])),
                          pack_bf16(__uint_as_float(l[1]), __uint_as_float(l[3])), pack_bf16(w[0], w[2]),
                          pack_bf16(w[1], w[3]), bc0, bc1);
#else
        mma_m16n8k8_tf32(d, h[0], h[1], h[2], h[3], bl0, bl1, d);
        mma_m16n8k8_tf32(d, l[0], l[1], l[2], l[3], bh0, bh1, d);
#endif
    };
    auto times = [&](const float (&d)[4], float2 b0, float2 b1, bool a0, bool a1) {   // w = d . b where live
        if (a0) { w[0] = d[0] * b0.x; w[2] = d[1] * b0.y; }
        if (a1) { w[1] = d[2] * b1.x; w[3] = d[3] * b1.y; }
    };
    auto rescale = [&](bool a0, bool a1) {                   // row sums of h -> pending power-of-two factors
        float s[4];
        mma_m16n8k8_tf32(s, h[0], h[1], h[2], h[3], kOne, kOne, zero4);
        // s >= 0, so the word shifted right by 23 IS the exponent field; 2^-e as one integer multiply-add (the integer
        // work sits on the half-rate ALU pipe: three instructions per factor instead of five)
        const uint32_t f0 = __float_as_uint(s[0]) >> 23, f1 = __float_as_uint(s[2]) >> 23;
        if (a0) { esum[0] += (int)f0 - 127; sc[0] = __uint_as_float(0x7f000000u - f0 * 0x800000u); }
        if (a1) { esum[1] += (int)f1 - 127; sc[1] = __uint_as_float(0x7f000000u - f1 * 0x800000u); }
    };

    if constexpr (VEC) {
        static_assert(DT == KHMM_F32 && kBlkM == 8, "vector path: float observations, 8-step blocks");
        // block k = recursion steps 8k .. 8k+7 = elements 8k .. 8k+7 forward, T-8-8k .. T-1-8k (reversed) backward
        const float *row[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int64_t b = tile * 16 + g + 8 * q;
            row[q] = (const float *)obs + (b < B ? b : tile * 16) * T;
        }
        auto load_vec = [&](float (&x)[2][kBlkM], int s0) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const float4 *p = (const float4 *)(row[q] + (DIR ? (int)T - 8 - s0 : s0));
                const float4 lo = __ldg(p), hi = __ldg(p + 1);
                const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                for (int i = 0; i < kBlkM; i++) x[q][i] = DIR ? v[7 - i] : v[i];
            }
        };
        float xs[2][kBlkM], xn[2][kBlkM];
        load_vec(xs, 0);
        load_vec(xn, 8);
        {   // step 0: start vector (pi forward, ones backward) times the first emission; steps 1..7 rescale at every step
            const float2 st = DIR ? live : pv;
            const float d0[4] = {st.x, st.y, st.x, st.y};
            times(d0, emis(xs[0][0]), emis(xs[1][0]), true, true);
            sc[0] = sc[1] = 1.f;
            split();
            rescale(true, true);
#pragma unroll
            for (int i = 1; i < kBlkM; i++) {
                float d[4];
                const float2 b0 = __fmul2_rn(emis(xs[0][i]), make_float2(sc[0], sc[0]));
                const float2 b1 = __fmul2_rn(emis(xs[1][i]), make_float2(sc[1], sc[1]));
                product(d);
                times(d, b0, b1, true, true);
                split();
                rescale(true, true);
            }
        }
        for (int s0 = 8; s0 < (int)T; s0 += kBlkM) {
#pragma unroll
            for (int q = 0; q < 2; q++)
#pragma unroll
                for (int i = 0; i < kBlkM; i++) xs[q][i] = xn[q][i];
            if (s0 + kBlkM < (int)T) load_vec(xn, s0 + kBlkM);   // the block after this one, in flight meanwhile
#pragma unroll
            for (int i = 0; i < kBlkM; i += 2) {
                float d[4];
                // scaled step (takes the pending factors), then a plain step whose row sums give the next factors
                const float2 b0 = __fmul2_rn(emis(xs[0][i]), make_float2(sc[0], sc[0]));
                const float2 b1 = __fmul2_rn(emis(xs[1][i]), make_float2(sc[1], sc[1]));
                product(d);
                times(d, b0, b1, true, true);
                split();
                const float2 c0 = emis(xs[0][i + 1]), c1 = emis(xs[1][i + 1]);
                product(d);
                times(d, c0, c1, true, true);
                split();
                rescale(true, true);
            }
        }
    } else {
    {   // step 0: start vector (pi forward, ones backward) times the first emission
        const float2 st = DIR ? live : pv;
        const float d0[4] = {st.x, st.y, st.x, st.y};
        times(d0, emis(load_x(0, 0)), emis(load_x(1, 0)), true, true);
        sc[0] = sc[1] = 1.f;
        split();
        rescale(true, true);
    }
    // Measured and dropped (config-2 size, 0.316 ms as below): alternating two observation buffers over a loop unrolled
    // twice to save the register copies between blocks (0.330 ms: the loop no longer sits well in the instruction
    // cache); one lane of a quad loading each observation and a quad shuffle handing it out, for a quarter of the L1 tag
    // look-ups (0.330 ms: the shuffle latency lands on the step); evaluating the emission values one step ahead by hand
    // (0.32 - 0.36 ms); 16-step or 4-step blocks (0.46 / 0.345 ms).
    float xs[2][kBlkM], xn[2][kBlkM];
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
        for (int i = 0; i < kBlkM; i++) xs[q][i] = load_x(q, 1 + i);
    for (int s0 = 1; s0 < lmax; s0 += kBlkM) {
        // the block after this one, in flight while this one is computed
        if (__all_sync(0xffffffffu, s0 + 2 * kBlkM <= len[0] && s0 + 2 * kBlkM <= len[1])) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int64_t at = first[q] + (DIR ? -(s0 + kBlkM) : (s0 + kBlkM));
#pragma unroll
                for (int i = 0; i < kBlkM; i++) xn[q][i] = obs_elem<DT>(obs, at + (DIR ? -i : i), M);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 2; q++)
#pragma unroll
                for (int i = 0; i < kBlkM; i++) xn[q][i] = load_x(q, s0 + kBlkM + i);
        }
        if (__all_sync(0xffffffffu, s0 + kBlkM <= len[0] && s0 + kBlkM <= len[1])) {
#pragma unroll
            for (int i = 0; i < kBlkM; i += 2) {
                float d[4];
                // scaled step (takes the pending factors), then a plain step whose row sums give the next factors
                const float2 b0 = __fmul2_rn(emis(xs[0][i]), make_float2(sc[0], sc[0]));
                const float2 b1 = __fmul2_rn(emis(xs[1][i]), make_float2(sc[1], sc[1]));
                product(d);
                times(d, b0, b1, true, true);
                split();
                const float2 c0 = emis(xs[0][i + 1]), c1 = emis(xs[1][i + 1]);
                product(d);
                times(d, c0, c1, true, true);
                split();
                rescale(true, true);
            }
        } else {
#pragma unroll
            for (int i = 0; i < kBlkM; i++) {                 // the ragged tail rescales at every step
                const bool a0 = s0 + i < len[0], a1 = s0 + i < len[1];
                float d[4];
                const float2 b0 = __fmul2_rn(emis(xs[0][i]), make_float2(sc[0], sc[0]));
                const float2 b1 = __fmul2_rn(emis(xs[1][i]), make_float2(sc[1], sc[1]));
                product(d);
                times(d, b0, b1, a0, a1);
                split();
                rescale(a0, a1);
            }
        }
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int i = 0; i < kBlkM; i++) xs[q][i] = xn[q][i];
    }
    }   // !VEC
    // closing dot product: forward sum(w), backward pi . w_0 -- of the vector as it stands (the pending factor was
    // counted in esum but not applied yet, so it is applied here)
    const float2 wf = DIR ? pv : live;
    float f[2] = {w[0] * wf.x + w[2] * wf.y, w[1] * wf.x + w[3] * wf.y};
#pragma unroll
    for (int q = 0; q < 2; q++) {
        f[q] += __shfl_xor_sync(0xffffffffu, f[q], 1);
        f[q] += __shfl_xor_sync(0xffffffffu, f[q], 2);
        const int64_t b = tile * 16 + g + 8 * q;
        if (c != 0 || b >= B) continue;
        const float fin = f[q] * sc[q];
        double ll = (double)esum[q] * 0.6931471805599453 + log((double)fin);
        if (!(fin > 0.f) || !(fin < INFINITY)) ll = __longlong_as_double(0x7ff8000000000000ll);  // underflow -> NaN (Q15)
        out[b] = ll;
    }
}

template <bool GAUSS, int DT, bool VEC>
__global__ void __launch_bounds__(64)
fwdbwd_mma8_kernel(int64_t B, int64_t T, int N, int M, const float *__restrict__ pi, const float *__restrict__ A,
                   const float *__restrict__ p0, const float *__restrict__ p1, const void *__restrict__ obs,
                   const int32_t *__restrict__ lengths, double *__restrict__ loglik, double *__restrict__ loglik_bwd) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const bool two = loglik_bwd != nullptr;
    const int64_t tile = two ? (warp >> 1) : warp;          // 16 sequences; even warps forward, odd warps backward
    if (tile * 16 >= B) return;
    if (two && (warp & 1))
        mma8_body<GAUSS, DT, 1, VEC>(B, T, N, M, pi, A, p0, p1, obs, lengths, loglik_bwd, tile);
    else
        mma8_body<GAUSS, DT, 0, VEC>(B, T, N, M, pi, A, p0, p1, obs, lengths, loglik, tile);
}

template <bool GAUSS, int DT>
static int launch_mma8_typed(int64_t B, int64_t T, int N, int M, const float *pi, const float *A, const float *p0,
                             const float *p1, const void *obs, const int32_t *lengths, double *loglik,
                             double *loglik_bwd, cudaStream_t st) {
    const int64_t tiles = (B + 15) / 16;
    const int64_t warps = tiles * (loglik_bwd ? 2 : 1);
    // whole sequences, blocks of 8 steps, 16-byte aligned rows -> the vector-load instantiation
    const bool vec = GAUSS && DT == KHMM_F32 && !lengths && T >= 16 && T % 8 == 0 && ((uintptr_t)obs & 15) == 0 &&
                     dev_switch(SW_SMALL_MMA) != 2;
    if constexpr (GAUSS && DT == KHMM_F32) {
        if (vec) {
            fwdbwd_mma8_kernel<GAUSS, DT, true><<<(unsigned)((warps + 1) / 2), 64, 0, st>>>(
                B, T, N, M, pi, A, p0, p1, obs, lengths, loglik, loglik_bwd);
            KHMM_LAUNCH_CHECK("fwdbwd_mma8_kernel");
            return KHMM_OK;
        }
    }
    fwdbwd_mma8_kernel<GAUSS, DT, false><<<(unsigned)((warps + 1) / 2), 64, 0, st>>>(B, T, N, M, pi, A, p0, p1, obs, lengths,
                                                                                   loglik, loglik_bwd);
    KHMM_LAUNCH_CHECK("fwdbwd_mma8_kernel");
    return KHMM_OK;
}

template <bool GAUSS>
static int launch_mma8(int64_t B, int64_t T, int N, int M, const float *pi, const float *A, const float *p0,
                       const float *p1, const void *obs, int obs_dtype, const int32_t *lengths, double *loglik,
                       double *loglik_bwd, cudaStream_t st) {
#define KHMM_MMA8(DT) launch_mma8_typed<GAUSS, DT>(B, T, N, M, pi, A, p0, p1, obs, lengths, loglik, loglik_bwd, st)
    if constexpr (GAUSS) {
        return obs_dtype == KHMM_F32 ? KHMM_MMA8(KHMM_F32) : KHMM_MMA8(KHMM_F64);
    } else {
        switch (obs_dtype) {
            case KHMM_U8: return KHMM_MMA8(KHMM_U8);
            case KHMM_U16: return KHMM_MMA8(KHMM_U16);
            case KHMM_I32: return KHMM_MMA8(KHMM_I32);
            default: return KHMM_MMA8(KHMM_I64);
        }
    }
#undef KHMM_MMA8
}

template <int NP, bool GAUSS, bool MX_SMEM>
static int launch_small(int64_t B, int64_t T, int N, int M, const float *pi, const float *A, const float *p0,
                        const float *p1, const void *obs, int obs_dtype, const int32_t *lengths, double *loglik,
                        double *loglik_bwd, cudaStream_t st) {
    const int dirs = loglik_bwd ? 2 : 1;
    const int threads = 64;  // one forward + one backward warp per CTA
    // Two sequences per thread (Q = 2) was measured SLOWER on B200 (0.47 ms vs 0.38 ms at N=8,
    // B=16384, T=2048): one warp with two interleaved chains does not out-issue two warps with one
    // chain each.  Kept as a template parameter for experiments, disabled in the dispatch.
    const bool q2 = false;
    int64_t warps = (B + (q2 ? 63 : 31)) / (q2 ? 64 : 32) * dirs;
    int64_t grid = (warps * 32 + threads - 1) / threads;
    if (q2)
        fwdbwd_small_kernel<NP, GAUSS, MX_SMEM, (NP <= 8 ? 2 : 1)><<<(unsigned)grid, threads, 0, st>>>(
            B, T, N, M, pi, A, p0, p1, obs, obs_dtype, lengths, loglik, loglik_bwd);
    else
        fwdbwd_small_kernel<NP, GAUSS, MX_SMEM, 1><<<(unsigned)grid, threads, 0, st>>>(
            B, T, N, M, pi, A, p0, p1, obs, obs_dtype, lengths, loglik, loglik_bwd);
    KHMM_LAUNCH_CHECK("fwdbwd_small_kernel");
    return KHMM_OK;
}

template <bool GAUSS>
static int dispatch_small(int64_t B, int64_t T, int N, int M, const float *pi, const float *A, const float *p0,
                          const float *p1, const void *obs, int obs_dtype, const int32_t *lengths, double *loglik,
                          double *loglik_bwd, khmm_stream_t stream) {
    KHMM_REQUIRE(B >= 0 && T >= 1 && N >= 1, "fwdbwd_loglik: bad shape");
    if (N > KHMM_SMALL_N_MAX) {
        set_error("fwdbwd_loglik: N=%d > %d; use forward/backward", N, KHMM_SMALL_N_MAX);
        return KHMM_ERR_UNSUPPORTED;
    }
    if (B == 0) return KHMM_OK;
    KHMM_REQUIRE(pi && A && p0 && obs && loglik, "fwdbwd_loglik: NULL pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 2) return launch_small<2, GAUSS, false>(B, T, N, M, pi, A, p0, p1, obs, obs_dtype, lengths, loglik, loglik_bwd, st);
    if (N <= 4) return launch_small<4, GAUSS, false>(B, T, N, M, pi, A, p0, p1, obs, obs_dtype, lengths, loglik, loglik_bwd, st);
    if (N > 4 && N <= 8 && dev_switch(SW_SMALL_MMA) != 0)
        return launch_mma8<GAUSS>(B, T, N, M, pi, A, p0, p1, obs, obs_dtype, lengths, loglik, loglik_bwd, st);
    if (N <= 8) return launch_small<8, GAUSS, false>(B, T, N, M, pi, A, p0, p1, obs, obs_dtype, lengths, loglik, loglik_bwd, st);
    return launch_small<16, GAUSS, true>(B, T, N, M, pi, A, p0, p1, obs, obs_dtype, lengths, loglik, loglik_bwd, st);
}

}  // namespace khmm

using namespace khmm;

extern "C" {

int khmm_fwdbwd_loglik_gauss_f32(int64_t B, int64_t T, int N, const float *pi, const float *A, const float *mean,
                                 const float *sd, const void *obs, int obs_dtype, const int32_t *lengths,
                                 double *loglik, double *loglik_bwd, khmm_stream_t stream) {
    KHMM_REQUIRE(sd, "fwdbwd_loglik_gauss: sd is NULL");
    KHMM_REQUIRE(obs_dtype == KHMM_F32 || obs_dtype == KHMM_F64, "gauss: obs dtype code %d is not f32/f64", obs_dtype);
    return dispatch_small<true>(B, T, N, 0, pi, A, mean, sd, obs, obs_dtype, lengths, loglik, loglik_bwd, stream);
}

int khmm_fwdbwd_loglik_discrete_f32(int64_t B, int64_t T, int N, int M, const float *pi, const float *A,
                                    const float *Bsm, const void *obs, int obs_dtype, const int32_t *lengths,
                                    double *loglik, double *loglik_bwd, khmm_stream_t stream) {
    KHMM_REQUIRE(M >= 1, "fwdbwd_loglik_discrete: M < 1");
    KHMM_REQUIRE(obs_dtype >= KHMM_U8 && obs_dtype <= KHMM_I64, "discrete: obs dtype code %d is not an integer type",
                 obs_dtype);
    return dispatch_small<false>(B, T, N, M, pi, A, Bsm, nullptr, obs, obs_dtype, lengths, loglik, loglik_bwd, stream);
}

}  // extern "C"
