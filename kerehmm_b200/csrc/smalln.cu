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

    // w M: input states in ONE accumulation chain per output pair (the four / eight chains of a step are independent,
    // which is all the instruction-level parallelism a warp needs with packed FFMA2 issuing every other cycle; the
    // round-1 version split even and odd inputs into two chains and paid a packed add per output pair to merge them)
    __device__ __forceinline__ void matvec(const float2 (&w)[H], float2 (&out)[H]) const {
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
