// Baum-Welch: xi accumulation (the GEMM-shaped part of AbstractHMM.zi, abstract_hmm.py:299-325,
// summed over t as discrete_hmm.py:155 does) and the device M-step (discrete_hmm.py:132-161 +
// util.smooth_probabilities, util.py:39-63).
//
// xi_raw[i][j] += sum over rows r=(b,t), t < len_b-1, of alpha[r][i] * V[r+1][j]
// (V[b][t][j] = b_j(o_t) beta_t(j) was written by the E-step backward kernel); the elementwise
// A[i][j] factor of zi is applied in the M-step.  zi is never materialised (67 MB per sequence at
// N=128, T=512 in the reference).  This is a (N x K) x (K x N) contraction with K = B*(T-1): both
// operands are K-major ("TN"), so tiles stream straight from the [B][T][N] stashes.
#include "common.cuh"

namespace khmm {

constexpr int XT = 128;      // tile edge (i and j)
constexpr int XK = 8;        // rows per stage
constexpr int XFLUSH = 128;  // stages between fp32 -> fp64 accumulator flushes

template <typename R>
__global__ void __launch_bounds__(256)
xi_accumulate_kernel(int64_t rows, int64_t T, int N, const int32_t *__restrict__ lengths, const R *__restrict__ alpha,
                     const R *__restrict__ V, double *__restrict__ xi, int64_t rows_per_cta) {
    __shared__ __align__(16) R Ps[2][XK][XT];
    __shared__ __align__(16) R Qs[2][XK][XT];
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * XT, j0 = blockIdx.y * XT;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 8 x 8 outputs each (strided by 16)
    const int64_t r_begin = (int64_t)blockIdx.z * rows_per_cta;
    const int64_t r_end = min(rows, r_begin + rows_per_cta);
    if (r_begin >= r_end) return;

    // loader mapping: XK rows x XT cols = 1024 elements, 4 per thread (one row, 4 consecutive cols)
    const int lr = tid >> 5;         // 0..7
    const int lc = (tid & 31) * 4;   // 0..124
    const bool vec = (N % 4 == 0) && sizeof(R) == 4;

    auto valid_row = [&](int64_t r) -> bool {
        if (r >= r_end) return false;
        int64_t b = r / T;
        int t = (int)(r - b * T);
        int len = lengths ? lengths[b] : (int)T;
        return t < len - 1;
    };
    auto load4 = [&](const R *base, int64_t r, int c0, bool ok, R (&out)[4]) {
#pragma unroll
        for (int q = 0; q < 4; q++) out[q] = 0;
        if (!ok) return;
        const R *p = base + r * N + c0;
        if (vec && c0 + 3 < N) {
            float4 f = *reinterpret_cast<const float4 *>(p);
            out[0] = (R)f.x; out[1] = (R)f.y; out[2] = (R)f.z; out[3] = (R)f.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (c0 + q < N) out[q] = p[q];
        }
    };

    double acc64[8][8];
    R acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) { acc64[a][b] = 0.0; acc[a][b] = 0; }

    R pr[4], qr[4];
    {
        int64_t r = r_begin + lr;
        bool ok = valid_row(r);
        load4(alpha, r, i0 + lc, ok, pr);
        load4(V, r + 1, j0 + lc, ok, qr);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) { Ps[0][lr][lc + q] = pr[q]; Qs[0][lr][lc + q] = qr[q]; }
    __syncthreads();

    int buf = 0, stage = 0;
    for (int64_t r0 = r_begin; r0 < r_end; r0 += XK, stage++) {
        int64_t rn = r0 + XK + lr;
        bool has_next = r0 + XK < r_end;
        if (has_next) {
            bool ok = valid_row(rn);
            load4(alpha, rn, i0 + lc, ok, pr);
            load4(V, rn + 1, j0 + lc, ok, qr);
        }
#pragma unroll
        for (int k = 0; k < XK; k++) {
            R p[8], q[8];
#pragma unroll
            for (int a = 0; a < 8; a++) p[a] = Ps[buf][k][ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 8; b++) q[b] = Qs[buf][k][tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int b = 0; b < 8; b++) acc[a][b] = fma(p[a], q[b], acc[a][b]);
        }
        if (sizeof(R) == 4 && (stage % XFLUSH) == XFLUSH - 1) {
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int b = 0; b < 8; b++) { acc64[a][b] += (double)acc[a][b]; acc[a][b] = 0; }
        }
        if (has_next) {
#pragma unroll
            for (int q = 0; q < 4; q++) { Ps[buf ^ 1][lr][lc + q] = pr[q]; Qs[buf ^ 1][lr][lc + q] = qr[q]; }
        }
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) {
            int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
            if (i < N && j < N) {
                double v = acc64[a][b] + (double)acc[a][b];
                if (v != 0.0) atomicAdd(xi + (int64_t)i * N + j, v);
            }
        }
}

template <typename R>
static int xi_impl(int64_t B, int64_t T, int N, const int32_t *lengths, const R *alpha, const R *V, double *xi,
                   khmm_stream_t stream) {
    KHMM_REQUIRE(B >= 0 && T >= 1 && N >= 1 && N <= 1024, "xi_accumulate: bad shape");
    if (B == 0 || T < 2) return KHMM_OK;
    KHMM_REQUIRE(alpha && V && xi, "xi_accumulate: NULL pointer argument");
    int64_t rows = B * T - 1;  // the last row has no successor
    int tiles = (N + XT - 1) / XT;
    int64_t want = ((int64_t)sm_count() * 2 + (int64_t)tiles * tiles - 1) / ((int64_t)tiles * tiles);
    int64_t max_split = (rows + XK * 16 - 1) / (XK * 16);
    int64_t nsplit = want < max_split ? want : max_split;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 65535) nsplit = 65535;
    int64_t rows_per = ((rows + nsplit - 1) / nsplit + XK - 1) / XK * XK;
    nsplit = (rows + rows_per - 1) / rows_per;
    dim3 grid(tiles, tiles, (unsigned)nsplit);
    xi_accumulate_kernel<R><<<grid, 256, 0, (cudaStream_t)stream>>>(rows, T, N, lengths, alpha, V, xi, rows_per);
    KHMM_LAUNCH_CHECK("xi_accumulate_kernel");
    return KHMM_OK;
}

// ----------------------------------------------------------------------------- M-step
__device__ __forceinline__ double block_sum_d(double v, double *red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect red from the previous use
    if (lane == 0) red[w] = v;
    __syncthreads();
    double x = 0;
    for (int k = 0; k < nw; k++) x += red[k];
    return x;
}

__device__ __forceinline__ bool isclose1(double s) { return fabs(s - 1.0) <= 1e-8 + 1e-5 * 1.0; }

// blocks [0, N): transition rows; [N, 2N): emission rows; 2N: pi.  row buffer lives in global out.
__global__ void __launch_bounds__(256)
mstep_discrete_kernel(int N, int M, const double *__restrict__ counts, const double *__restrict__ A,
                      double *__restrict__ pi_new, double *__restrict__ A_new, double *__restrict__ B_new,
                      int32_t *__restrict__ status) {
    __shared__ double red[32];
    const double *c_pi = counts;
    const double *c_xi = c_pi + N;
    const double *c_gam = c_xi + (int64_t)N * N;
    const double *c_num = c_gam + N;
    const double *c_den = c_num + (int64_t)M * N;
    const double *c_tail = c_den + N;
    const double DP = 1e-10;
    int blk = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    int st = 0;
    if (blk == 2 * N) {
        double nseq = c_tail[1];
        double s = 0;
        for (int i = tid; i < N; i += nt) {
            double v = c_pi[i] / nseq;
            pi_new[i] = v;
            s += v;
        }
        s = block_sum_d(s, red);
        if (!isclose1(s)) st |= 4;
        if (c_tail[2] != 0.0) st |= 1;
    } else if (blk >= N) {
        // emission row: B_new[i] = smooth1d(num[i]/den[i])   (discrete_hmm.py:137-146, util.py:56-60)
        int i = blk - N;
        double *row = B_new + (int64_t)i * M;
        double den = c_den[i];
        double cnt = 0;
        for (int k = tid; k < M; k += nt) {
            double v = c_num[(int64_t)k * N + i] / den;
            row[k] = v;
            if (v < DP) cnt += 1;
        }
        cnt = block_sum_d(cnt, red);
        int kz = (int)cnt;
        if (kz >= M) {
            st |= 1;
        } else {
            double sub = (double)kz * DP / (double)(M - kz);
            double s = 0;
            for (int k = tid; k < M; k += nt) {
                double v = row[k] - sub;
                if (v < DP) v = DP;
                row[k] = v;
                s += v;
            }
            s = block_sum_d(s, red);
            double s2 = 0;
            for (int k = tid; k < M; k += nt) {
                double v = row[k] / s;
                row[k] = v;
                s2 += v;
            }
            s2 = block_sum_d(s2, red);
            if (!isclose1(s2)) st |= 4;
        }
    } else {
        // transition row: xi/gam, row-normalise, smooth2d (discrete_hmm.py:155-161, util.py:46-55)
        int i = blk;
        double *row = A_new + (int64_t)i * N;
        double gam = c_gam[i];
        double s = 0;
        for (int j = tid; j < N; j += nt) {
            double v = c_xi[(int64_t)i * N + j] * A[(int64_t)i * N + j] / gam;
            row[j] = v;
            s += v;
        }
        s = block_sum_d(s, red);
        double s1 = 0;
        for (int j = tid; j < N; j += nt) {
            double v = row[j] / s;
            row[j] = v;
            s1 += v;
        }
        s1 = block_sum_d(s1, red);
        double cnt = 0, zsum = 0;
        for (int j = tid; j < N; j += nt) {
            double v = row[j] / s1;
            row[j] = v;
            if (v < DP) { cnt += 1; zsum += v; }
        }
        cnt = block_sum_d(cnt, red);
        zsum = block_sum_d(zsum, red);
        int kz = (int)cnt;
        if (kz > 0) {
            if (kz == 1 || kz == N) {
                // util.py:51-53: the "counter" is the array of small values z; k == 1 broadcasts the
                // single z over the row, k == n subtracts elementwise.
                for (int j = tid; j < N; j += nt) {
                    double z = (kz == 1) ? zsum : row[j];
                    double v = row[j] - z * DP / ((double)N - z);
                    if (v < DP) v = DP;
                    row[j] = v;
                }
            } else {
                st |= 2;  // ValueError in the reference (operands could not be broadcast)
            }
        }
        double s2 = 0;
        for (int j = tid; j < N; j += nt) s2 += row[j];
        s2 = block_sum_d(s2, red);
        double s3 = 0;
        for (int j = tid; j < N; j += nt) {
            double v = row[j] / s2;
            row[j] = v;
            s3 += v;
        }
        s3 = block_sum_d(s3, red);
        if (!isclose1(s3)) st |= 4;
    }
    if (tid == 0 && st) atomicOr(status, st);
}

}  // namespace khmm

using namespace khmm;

extern "C" {

int khmm_xi_accumulate_f32(int64_t B, int64_t T, int N, const int32_t *lengths, const float *alpha, const float *V,
                           double *xi_raw, khmm_stream_t stream) {
    return xi_impl<float>(B, T, N, lengths, alpha, V, xi_raw, stream);
}
int khmm_xi_accumulate_f64(int64_t B, int64_t T, int N, const int32_t *lengths, const double *alpha, const double *V,
                           double *xi_raw, khmm_stream_t stream) {
    return xi_impl<double>(B, T, N, lengths, alpha, V, xi_raw, stream);
}

int khmm_mstep_discrete_f64(int N, int M, const double *counts, const double *A, double *pi_new, double *A_new,
                            double *B_new, int32_t *status, khmm_stream_t stream) {
    KHMM_REQUIRE(N >= 1 && M >= 1 && counts && A && pi_new && A_new && B_new && status, "mstep: bad arguments");
    KHMM_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int32_t), (cudaStream_t)stream));
    mstep_discrete_kernel<<<2 * N + 1, 256, 0, (cudaStream_t)stream>>>(N, M, counts, A, pi_new, A_new, B_new, status);
    KHMM_LAUNCH_CHECK("mstep_discrete_kernel");
    return KHMM_OK;
}

}  // extern "C"
