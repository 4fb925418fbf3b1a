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
// Scaling uses the hardware reciprocal r_t ~ 1/c_t and accumulates log P from r_t itself
// (exponent in an integer, mantissa product in a double), which makes the result independent of
// the reciprocal's rounding and costs no transcendental per step.
#include "common.cuh"

namespace khmm {

template <int NP>
struct SmallConsts {  // per-thread (identical across threads) model constants, in registers
    float mx[NP][NP];  // mx[y][x]: out[x] = sum_y in[y]*mx[y][x]  (A forward, A^T backward)
    float e0[NP], e1[NP], e2[NP];  // gaussian: mean, k=-0.5*log2(e)/sd^2, log2(1/(sd*sqrt(2pi)))
};

struct LogAccum {  // log2 of a product of positive floats, exact to double rounding
    int esum = 0, n = 0;
    double mprod = 1.0, l2 = 0.0;
    bool bad = false;
    __device__ __forceinline__ void mul(float r) {
        unsigned bits = __float_as_uint(r);
        int e = (bits >> 23) & 0xff;
        bad |= (e == 0) | (e == 255) | ((bits >> 31) != 0);
        esum += e;
        n++;
        mprod *= (double)__uint_as_float((bits & 0x7fffffu) | 0x3f800000u);
        if ((n & 255) == 0) { l2 += log2(mprod); mprod = 1.0; }
    }
    __device__ __forceinline__ double log2_total() const {
        if (bad) return __longlong_as_double(0x7ff8000000000000ll);
        return (double)(esum - 127 * n) + l2 + log2(mprod);
    }
};

template <int NP, bool GAUSS>
__device__ __forceinline__ void emis_small(const SmallConsts<NP> &mc, float x, int sym, const float *__restrict__ Bsm,
                                           int N, float (&bv)[NP]) {
    if (GAUSS) {
#pragma unroll
        for (int j = 0; j < NP; j++) {
            float d = x - mc.e0[j];
            bv[j] = exp2f(fmaf(d * d, mc.e1[j], mc.e2[j]));
        }
    } else {
        const float *row = Bsm + (int64_t)sym * N;
#pragma unroll
        for (int j = 0; j < NP; j++) bv[j] = j < N ? __ldg(row + j) : 0.f;
    }
}

template <int NP, bool GAUSS>
__global__ void __launch_bounds__(128)
fwdbwd_small_kernel(int64_t B, int64_t T, int N, int M, const float *__restrict__ pi, const float *__restrict__ A,
                    const float *__restrict__ p0, const float *__restrict__ p1, const void *__restrict__ obs,
                    int obs_dtype, const int32_t *__restrict__ lengths, double *__restrict__ loglik,
                    double *__restrict__ loglik_bwd) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const bool two = loglik_bwd != nullptr;
    const int dir = two ? (int)(warp & 1) : 0;  // warp-uniform
    const int64_t b = (two ? (warp >> 1) : warp) * 32 + lane;

    SmallConsts<NP> mc;
#pragma unroll
    for (int y = 0; y < NP; y++)
#pragma unroll
        for (int x = 0; x < NP; x++) {
            float v = 0.f;
            if (y < N && x < N) v = dir ? __ldg(A + x * N + y) : __ldg(A + y * N + x);
            mc.mx[y][x] = v;
        }
    float pv[NP];
#pragma unroll
    for (int j = 0; j < NP; j++) {
        pv[j] = j < N ? __ldg(pi + j) : 0.f;
        if (GAUSS) {
            float m = j < N ? __ldg(p0 + j) : 0.f, sd = j < N ? __ldg(p1 + j) : 1.f;
            mc.e0[j] = m;
            mc.e1[j] = j < N ? -0.5f * 1.4426950408889634f / (sd * sd) : 0.f;
            mc.e2[j] = j < N ? -log2f(sd * 2.5066282746310002f) : -INFINITY;
        }
    }
    if (b >= B) return;
    const int len = lengths ? lengths[b] : (int)T;
    const int64_t base = b * T;

    float u[NP];
    LogAccum acc;
    for (int step = 0; step < len; step++) {
        const int t = dir ? len - 1 - step : step;
        float x = 0.f;
        int sym = 0;
        if (GAUSS) {
            x = load_real<float>(obs, obs_dtype, base + t);
        } else {
            sym = load_symbol(obs, obs_dtype, base + t);
            sym = sym < 0 ? 0 : (sym >= M ? M - 1 : sym);
        }
        float bv[NP];
        emis_small<NP, GAUSS>(mc, x, sym, p0, N, bv);
        float val[NP];
        if (step == 0) {
#pragma unroll
            for (int j = 0; j < NP; j++) val[j] = (dir ? (j < N ? 1.f : 0.f) : pv[j]) * bv[j];
        } else {
#pragma unroll
            for (int xx = 0; xx < NP; xx++) {
                float s = 0.f;
#pragma unroll
                for (int y = 0; y < NP; y++) s = fmaf(u[y], mc.mx[y][xx], s);
                val[xx] = s * bv[xx];
            }
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NP; j++) s += val[j];
        float r = __frcp_rn(s);
        acc.mul(r);
#pragma unroll
        for (int j = 0; j < NP; j++) u[j] = val[j] * r;
    }
    // closing dot product: forward sum(u) (= 1 up to rounding), backward pi . w_0
    float fin = 0.f;
#pragma unroll
    for (int j = 0; j < NP; j++) fin = fmaf(u[j], dir ? pv[j] : (j < N ? 1.f : 0.f), fin);
    double ll = (-acc.log2_total()) * 0.6931471805599453 + log((double)fin);
    (dir ? loglik_bwd : loglik)[b] = ll;
}

template <int NP, bool GAUSS>
static int launch_small(int64_t B, int64_t T, int N, int M, const float *pi, const float *A, const float *p0,
                        const float *p1, const void *obs, int obs_dtype, const int32_t *lengths, double *loglik,
                        double *loglik_bwd, cudaStream_t st) {
    int64_t warps = (B + 31) / 32 * (loglik_bwd ? 2 : 1);
    int threads = 64;  // one forward + one backward warp per CTA
    int64_t grid = (warps * 32 + threads - 1) / threads;
    fwdbwd_small_kernel<NP, GAUSS><<<(unsigned)grid, threads, 0, st>>>(B, T, N, M, pi, A, p0, p1, obs, obs_dtype,
                                                                      lengths, loglik, loglik_bwd);
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
    KHMM_REQUIRE(pi && A && p0 && obs && loglik, "fwdbwd_loglik: NULL pointer argument");
    if (B == 0) return KHMM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 4) return launch_small<4, GAUSS>(B, T, N, M, pi, A, p0, p1, obs, obs_dtype, lengths, loglik, loglik_bwd, st);
    if (N <= 8) return launch_small<8, GAUSS>(B, T, N, M, pi, A, p0, p1, obs, obs_dtype, lengths, loglik, loglik_bwd, st);
    return launch_small<16, GAUSS>(B, T, N, M, pi, A, p0, p1, obs, obs_dtype, lengths, loglik, loglik_bwd, st);
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
