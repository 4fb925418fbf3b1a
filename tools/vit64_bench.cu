// Experiment harness for the N=64 float64 Viterbi step (development tool, not product code):
// thread j keeps column j of logA in registers; several formulations of the first-index
// argmax update are timed against each other and checked against a plain reference kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/vit64_bench tools/vit64_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

constexpr int N = 64;

// ------------------------------------------------------------ reference (simple, trusted)
__global__ void ref_kernel(int B, int T, int M, const double *logpi, const double *logA, const double *logB,
                           const uint8_t *obs, uint8_t *psi, double *delta_out) {
    __shared__ double d[2][N];
    int b = blockIdx.x, j = threadIdx.x;
    d[0][j] = logpi[j] + logB[obs[(size_t)b * T] * N + j];
    __syncthreads();
    int cur = 0;
    for (int t = 1; t < T; t++) {
        double best = d[cur][0] + logA[j];
        int arg = 0;
        for (int i = 1; i < N; i++) {
            double v = d[cur][i] + logA[i * N + j];
            if (v > best) { best = v; arg = i; }
        }
        d[cur ^ 1][j] = best + logB[obs[(size_t)b * T + t] * N + j];
        psi[((size_t)b * T + t) * N + j] = (uint8_t)arg;
        __syncthreads();
        cur ^= 1;
    }
    delta_out[(size_t)b * N + j] = d[cur][j];
}

// ----------------------------------------------------------------- candidate kernels
// VARIANT 0: fp64 compare (DSETP) + selects
// VARIANT 1: integer compare of the bit patterns (valid because every value is <= +0.0)
// VARIANT 2: fp64 compare, value kept via predicated re-add
template <int VARIANT>
__device__ __forceinline__ void upd(double v, int i, double &best, int &arg) {
    if (VARIANT == 0) {
        if (v > best) { best = v; arg = i; }
    } else if (VARIANT == 1) {
        unsigned long long kv = (unsigned long long)__double_as_longlong(v);
        unsigned long long kb = (unsigned long long)__double_as_longlong(best);
        if (kv < kb) { best = v; arg = i; }
    } else if (VARIANT == 2) {
        bool p = v > best;
        best = p ? v : best;
        arg = p ? i : arg;
    } else if (VARIANT == 3) {  // compare on the fp64 pipe, value kept by a predicated fp64 add, index by a predicated mov
        asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %2, %0;\n\t@p add.f64 %0, %2, 0d0000000000000000;\n\t@p mov.s32 %1, %3;\n\t}"
            : "+d"(best), "+r"(arg) : "d"(v), "r"(i));
    } else if (VARIANT == 4) {  // analysis only: max without the index
        best = fmax(best, v);
    } else if (VARIANT == 5) {  // analysis only: adds only
        best += v;
    } else if (VARIANT == 6) {  // predicated moves of the two halves
        int lo = __double2loint(best), hi = __double2hiint(best);
        asm("{\n\t.reg .pred p;\n\t.reg .b32 vl, vh;\n\tsetp.gt.f64 p, %3, %4;\n\tmov.b64 {vl, vh}, %3;\n\t@p mov.b32 %0, vl;\n\t@p mov.b32 %1, vh;\n\t@p mov.s32 %2, %5;\n\t}"
            : "+r"(lo), "+r"(hi), "+r"(arg) : "d"(v), "d"(best), "r"(i));
        best = __hiloint2double(hi, lo);
    }
}

template <int S, int VARIANT, int CHAINS>
__global__ void __launch_bounds__(64)
vit64_kernel(int B, int T, int M, const double *__restrict__ logpi, const double *__restrict__ logA,
             const double *__restrict__ logB, const uint8_t *__restrict__ obs, uint8_t *__restrict__ psi,
             double *__restrict__ delta_out) {
    __shared__ __align__(16) double d[2][S][N];
    const int j = threadIdx.x;
    double a[N];
#pragma unroll
    for (int i = 0; i < N; i++) a[i] = logA[i * N + j];
    const double lp = logpi[j];
    for (int g = blockIdx.x; g * S < B; g += gridDim.x) {
        const int b0 = g * S;
#pragma unroll
        for (int s = 0; s < S; s++) {
            int b = min(b0 + s, B - 1);
            d[0][s][j] = lp + logB[obs[(size_t)b * T] * N + j];
        }
        __syncthreads();
        int cur = 0;
        for (int t = 1; t < T; t++) {
            double e[S];
#pragma unroll
            for (int s = 0; s < S; s++) {
                int b = min(b0 + s, B - 1);
                e[s] = __ldg(logB + obs[(size_t)b * T + t] * N + j);
            }
#pragma unroll
            for (int s = 0; s < S; s++) {
                double best[CHAINS];
                int arg[CHAINS];
                const double *dp = d[cur][s];
#pragma unroll
                for (int c = 0; c < CHAINS; c++) {
                    best[c] = dp[c * (N / CHAINS)] + a[c * (N / CHAINS)];
                    arg[c] = c * (N / CHAINS);
                }
#pragma unroll
                for (int k = 1; k < N / CHAINS; k++) {
#pragma unroll
                    for (int c = 0; c < CHAINS; c++) {
                        int i = c * (N / CHAINS) + k;
                        upd<VARIANT>(dp[i] + a[i], i, best[c], arg[c]);
                    }
                }
                // merge chains in index order: a later chain wins only when strictly greater
#pragma unroll
                for (int c = 1; c < CHAINS; c++)
                    if (best[c] > best[0]) { best[0] = best[c]; arg[0] = arg[c]; }
                int b = b0 + s;
                if (b < B) psi[((size_t)b * T + t) * N + j] = (uint8_t)arg[0];
                d[cur ^ 1][s][j] = best[0] + e[s];
            }
            __syncthreads();
            cur ^= 1;
        }
#pragma unroll
        for (int s = 0; s < S; s++)
            if (b0 + s < B) delta_out[(size_t)(b0 + s) * N + j] = d[cur][s][j];
        __syncthreads();
    }
}

// ---------------------------------------------------------------- thread-per-sequence design
// delta of one sequence lives in the registers of ONE thread; logA comes through the constant
// bank (warp-uniform operand), so the inner step needs no shared-memory or register-file bandwidth
// beyond its own operands.
__constant__ double cAT[N * N];  // cAT[j*N + i] = logA[i][j]

template <int CHAINS, int JU, int TPB>
__global__ void __launch_bounds__(TPB)
vit64_tps_kernel(int B, int T, int M, const double *__restrict__ logpi, const double *__restrict__ logB,
                 const uint8_t *__restrict__ obs, uint8_t *__restrict__ psi, double *__restrict__ delta_out) {
    __shared__ double tmp[N][TPB];
    const int b = blockIdx.x * TPB + threadIdx.x;
    if (b >= B) return;
    const uint8_t *ob = obs + (size_t)b * T;
    double d[N];
    {
        const double *e = logB + (size_t)ob[0] * N;
#pragma unroll
        for (int i = 0; i < N; i++) d[i] = logpi[i] + e[i];
    }
    uint8_t *prow = psi + (size_t)b * T * N;
    for (int t = 1; t < T; t++) {
        const double *e = logB + (size_t)ob[t] * N;
        uint32_t pack = 0;
        uint32_t packs[4];
#pragma unroll 1
        for (int j0 = 0; j0 < N; j0 += JU) {
#pragma unroll
            for (int jj = 0; jj < JU; jj++) {
                const int j = j0 + jj;
                const double *col = cAT + j * N;
                double best[CHAINS];
                int arg[CHAINS];
#pragma unroll
                for (int c = 0; c < CHAINS; c++) {
                    best[c] = d[c * (N / CHAINS)] + col[c * (N / CHAINS)];
                    arg[c] = c * (N / CHAINS);
                }
#pragma unroll
                for (int k = 1; k < N / CHAINS; k++) {
#pragma unroll
                    for (int c = 0; c < CHAINS; c++) {
                        int i = c * (N / CHAINS) + k;
                        double v = d[i] + col[i];
                        if (v > best[c]) { best[c] = v; arg[c] = i; }
                    }
                }
#pragma unroll
                for (int c = 1; c < CHAINS; c++)
                    if (best[c] > best[0]) { best[0] = best[c]; arg[0] = arg[c]; }
                tmp[j][threadIdx.x] = best[0] + __ldg(e + j);
                pack |= (uint32_t)arg[0] << (8 * (j & 3));
                if ((j & 3) == 3) {
                    packs[(j >> 2) & 3] = pack;
                    pack = 0;
                    if ((j & 15) == 15)
                        *reinterpret_cast<uint4 *>(prow + (size_t)t * N + (j - 15)) =
                            make_uint4(packs[0], packs[1], packs[2], packs[3]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < N; i++) d[i] = tmp[i][threadIdx.x];
    }
#pragma unroll
    for (int i = 0; i < N; i++) delta_out[(size_t)b * N + i] = d[i];
}

// ------------------------------------------------------- 2 targets x 32 sources per thread
// CTA = 64 threads = 2 warps.  Warp w owns the sources [32w, 32w+32); lane l owns the targets l and
// l+32 and keeps the 2x32 tile of logA in registers.  Every delta value read from shared memory
// (a pure broadcast) now feeds two transitions, halving the shared->register traffic that bounds
// the one-target-per-thread layout.  The two warps exchange partial (best, arg) through shared
// memory once per (sequence, step); warp w finalises targets l+32w.
template <int S, int CH>
__global__ void __launch_bounds__(64)
vit64_t2_kernel(int B, int T, int M, const double *__restrict__ logpi, const double *__restrict__ logA,
                const double *__restrict__ logB, const uint8_t *__restrict__ obs, uint8_t *__restrict__ psi,
                double *__restrict__ delta_out) {
    __shared__ __align__(16) double d[2][S][N];
    __shared__ double pbest[S][N];   // partial best of the NON-finalising warp, per target
    __shared__ int parg[S][N];
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    double a0[32], a1[32];   // logA[32w + k][l], logA[32w + k][l + 32]
#pragma unroll
    for (int k = 0; k < 32; k++) {
        a0[k] = logA[(32 * w + k) * N + l];
        a1[k] = logA[(32 * w + k) * N + l + 32];
    }
    const int jfin = l + 32 * w;          // target this thread finalises
    const int joth = l + 32 * (1 - w);    // target whose partial it hands over
    const double lp = logpi[jfin];
    for (int g = blockIdx.x; g * S < B; g += gridDim.x) {
        const int b0 = g * S;
#pragma unroll
        for (int s = 0; s < S; s++) {
            int b = min(b0 + s, B - 1);
            d[0][s][jfin] = lp + logB[obs[(size_t)b * T] * N + jfin];
        }
        __syncthreads();
        int cur = 0;
        for (int t = 1; t < T; t++) {
            double e[S];
#pragma unroll
            for (int s = 0; s < S; s++) {
                int b = min(b0 + s, B - 1);
                e[s] = __ldg(logB + obs[(size_t)b * T + t] * N + jfin);
            }
            double fb[S];
            int fa[S];
#pragma unroll
            for (int s = 0; s < S; s++) {
                const double *dp = d[cur][s] + 32 * w;
                double best0[CH], best1[CH];
                int arg0[CH], arg1[CH];
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    int k = c * (32 / CH);
                    double dv = dp[k];
                    best0[c] = dv + a0[k]; arg0[c] = k;
                    best1[c] = dv + a1[k]; arg1[c] = k;
                }
#pragma unroll
                for (int kk = 1; kk < 32 / CH; kk++) {
#pragma unroll
                    for (int c = 0; c < CH; c++) {
                        int k = c * (32 / CH) + kk;
                        double dv = dp[k];
                        double v0 = dv + a0[k], v1 = dv + a1[k];
                        if (v0 > best0[c]) { best0[c] = v0; arg0[c] = k; }
                        if (v1 > best1[c]) { best1[c] = v1; arg1[c] = k; }
                    }
                }
#pragma unroll
                for (int c = 1; c < CH; c++) {
                    if (best0[c] > best0[0]) { best0[0] = best0[c]; arg0[0] = arg0[c]; }
                    if (best1[c] > best1[0]) { best1[0] = best1[c]; arg1[0] = arg1[c]; }
                }
                // own target: l + 32w  -> index 0 if w == 0 else 1
                double ob_ = w ? best1[0] : best0[0];
                int oa_ = (w ? arg1[0] : arg0[0]) + 32 * w;
                pbest[s][joth] = w ? best0[0] : best1[0];
                parg[s][joth] = (w ? arg0[0] : arg1[0]) + 32 * w;
                fb[s] = ob_;
                fa[s] = oa_;
            }
            __syncthreads();
#pragma unroll
            for (int s = 0; s < S; s++) {
                double qb = pbest[s][jfin];
                int qa = parg[s][jfin];
                // the other warp's sources are the upper half when w == 0 (it wins only if strictly
                // greater) and the lower half when w == 1 (it also wins ties)
                bool take = w ? (qb >= fb[s]) : (qb > fb[s]);
                double bb = take ? qb : fb[s];
                int ba = take ? qa : fa[s];
                int b = b0 + s;
                if (b < B) psi[((size_t)b * T + t) * N + jfin] = (uint8_t)ba;
                d[cur ^ 1][s][jfin] = bb + e[s];
            }
            __syncthreads();
            cur ^= 1;
        }
#pragma unroll
        for (int s = 0; s < S; s++)
            if (b0 + s < B) delta_out[(size_t)(b0 + s) * N + jfin] = d[cur][s][jfin];
        __syncthreads();
    }
}

// ------------------------------------------- 2 targets x (64/W) sources per thread, W warps
// Same idea as vit64_t2_kernel with W = 2 or 4 warps per CTA: warp w owns sources [w*SRC, (w+1)*SRC),
// lane l owns targets l and l+32 (a 2 x SRC register tile of logA).  Partials of all warps go
// through shared memory; threads 0..63 merge them in source order (first index wins ties).
template <int S, int W, int CH>
__global__ void __launch_bounds__(32 * W)
vit64_tw_kernel(int B, int T, int M, const double *__restrict__ logpi, const double *__restrict__ logA,
                const double *__restrict__ logB, const uint8_t *__restrict__ obs, uint8_t *__restrict__ psi,
                double *__restrict__ delta_out) {
    constexpr int SRC = N / W;
    __shared__ __align__(16) double d[2][S][N];
    __shared__ double pbest[S][W][N];
    __shared__ int parg[S][W][N];
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5, tid = threadIdx.x;
    double a0[SRC], a1[SRC];
#pragma unroll
    for (int k = 0; k < SRC; k++) {
        a0[k] = logA[(SRC * w + k) * N + l];
        a1[k] = logA[(SRC * w + k) * N + l + 32];
    }
    const double lp = tid < N ? logpi[tid] : 0.0;
    for (int g = blockIdx.x; g * S < B; g += gridDim.x) {
        const int b0 = g * S;
        if (tid < N) {
#pragma unroll
            for (int s = 0; s < S; s++) {
                int b = min(b0 + s, B - 1);
                d[0][s][tid] = lp + logB[obs[(size_t)b * T] * N + tid];
            }
        }
        __syncthreads();
        int cur = 0;
        for (int t = 1; t < T; t++) {
            double e[S];
            if (tid < N) {
#pragma unroll
                for (int s = 0; s < S; s++) {
                    int b = min(b0 + s, B - 1);
                    e[s] = __ldg(logB + obs[(size_t)b * T + t] * N + tid);
                }
            }
#pragma unroll
            for (int s = 0; s < S; s++) {
                const double *dp = d[cur][s] + SRC * w;
                double best0[CH], best1[CH];
                int arg0[CH], arg1[CH];
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    int k = c * (SRC / CH);
                    double dv = dp[k];
                    best0[c] = dv + a0[k]; arg0[c] = k;
                    best1[c] = dv + a1[k]; arg1[c] = k;
                }
#pragma unroll
                for (int kk = 1; kk < SRC / CH; kk++) {
#pragma unroll
                    for (int c = 0; c < CH; c++) {
                        int k = c * (SRC / CH) + kk;
                        double dv = dp[k];
                        double v0 = dv + a0[k], v1 = dv + a1[k];
                        if (v0 > best0[c]) { best0[c] = v0; arg0[c] = k; }
                        if (v1 > best1[c]) { best1[c] = v1; arg1[c] = k; }
                    }
                }
#pragma unroll
                for (int c = 1; c < CH; c++) {
                    if (best0[c] > best0[0]) { best0[0] = best0[c]; arg0[0] = arg0[c]; }
                    if (best1[c] > best1[0]) { best1[0] = best1[c]; arg1[0] = arg1[c]; }
                }
                pbest[s][w][l] = best0[0];      parg[s][w][l] = arg0[0] + SRC * w;
                pbest[s][w][l + 32] = best1[0]; parg[s][w][l + 32] = arg1[0] + SRC * w;
            }
            __syncthreads();
            if (tid < N) {
#pragma unroll
                for (int s = 0; s < S; s++) {
                    double bb = pbest[s][0][tid];
                    int ba = parg[s][0][tid];
#pragma unroll
                    for (int q = 1; q < W; q++) {
                        double qb = pbest[s][q][tid];
                        int qa = parg[s][q][tid];
                        if (qb > bb) { bb = qb; ba = qa; }
                    }
                    int b = b0 + s;
                    if (b < B) psi[((size_t)b * T + t) * N + tid] = (uint8_t)ba;
                    d[cur ^ 1][s][tid] = bb + e[s];
                }
            }
            __syncthreads();
            cur ^= 1;
        }
        if (tid < N) {
#pragma unroll
            for (int s = 0; s < S; s++)
                if (b0 + s < B) delta_out[(size_t)(b0 + s) * N + tid] = d[cur][s][tid];
        }
        __syncthreads();
    }
}

template <class F>
static float time_ms(F f, int reps = 3) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv) {
    int B = argc > 1 ? atoi(argv[1]) : 8192, T = argc > 2 ? atoi(argv[2]) : 512, M = 256;
    std::vector<double> hpi(N), hA(N * N), hB(M * N);
    std::vector<uint8_t> hobs((size_t)B * T);
    srand(1);
    auto rnd = [] { return (rand() + 1.0) / (RAND_MAX + 2.0); };
    for (auto &x : hpi) x = log(rnd() / N);
    for (auto &x : hA) x = log(rnd() / N);
    for (auto &x : hB) x = log(rnd() / M);
    for (auto &x : hobs) x = rand() % M;
    double *pi, *A, *Bm, *dref, *dout;
    uint8_t *obs, *psiref, *psi;
    cudaMalloc(&pi, N * 8); cudaMalloc(&A, N * N * 8); cudaMalloc(&Bm, M * N * 8);
    cudaMalloc(&obs, (size_t)B * T); cudaMalloc(&psiref, (size_t)B * T * N); cudaMalloc(&psi, (size_t)B * T * N);
    cudaMalloc(&dref, (size_t)B * N * 8); cudaMalloc(&dout, (size_t)B * N * 8);
    cudaMemcpy(pi, hpi.data(), N * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(A, hA.data(), N * N * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(Bm, hB.data(), M * N * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(obs, hobs.data(), (size_t)B * T, cudaMemcpyHostToDevice);
    cudaMemset(psiref, 0, (size_t)B * T * N);
    float tr = time_ms([&] { ref_kernel<<<B, N>>>(B, T, M, pi, A, Bm, obs, psiref, dref); }, 1);
    double trans = (double)B * T * N * N;
    printf("ref: %.2f ms  %.3e tr/s\n", tr, trans / tr * 1e3);
    std::vector<uint8_t> hp((size_t)B * T * N), hr((size_t)B * T * N);
    std::vector<double> hd((size_t)B * N), hdr((size_t)B * N);
    cudaMemcpy(hr.data(), psiref, hr.size(), cudaMemcpyDeviceToHost);
    cudaMemcpy(hdr.data(), dref, hdr.size() * 8, cudaMemcpyDeviceToHost);
    int sms = 148;
    auto run = [&](const char *name, auto kern, int S, int ctas_per_sm) {
        cudaMemset(psi, 0, (size_t)B * T * N);
        int grid = sms * ctas_per_sm;
        if (grid * S > B) grid = (B + S - 1) / S;
        float t = time_ms([&] { kern<<<grid, 64>>>(B, T, M, pi, A, Bm, obs, psi, dout); });
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hp.data(), psi, hp.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(hd.data(), dout, hd.size() * 8, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (size_t k = 0; k < hp.size(); k++) bad += hp[k] != hr[k];
        size_t badd = 0;
        for (size_t k = 0; k < hd.size(); k++) badd += hd[k] != hdr[k];
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 64, 0);
        printf("%-28s %.2f ms  %.3e tr/s  psi_mismatch=%zu delta_mismatch=%zu occ=%d %s\n", name, t, trans / t * 1e3,
               bad, badd, occ, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    run("S2 V0 C4", vit64_kernel<2, 0, 4>, 2, 4);
    run("S2 V1 C4", vit64_kernel<2, 1, 4>, 2, 4);
    run("S2 V2 C4", vit64_kernel<2, 2, 4>, 2, 4);
    run("S4 V0 C4", vit64_kernel<4, 0, 4>, 4, 4);
    run("S4 V0 C2", vit64_kernel<4, 0, 2>, 4, 4);
    run("S4 V0 C8", vit64_kernel<4, 0, 8>, 4, 4);
    run("S4 V1 C4", vit64_kernel<4, 1, 4>, 4, 4);
    {
        std::vector<double> hAT(N * N);
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) hAT[j * N + i] = hA[i * N + j];
        cudaMemcpyToSymbol(cAT, hAT.data(), N * N * 8);
    }
    auto run_tps = [&](const char *name, auto kern, int tpb) {
        cudaMemset(psi, 0, (size_t)B * T * N);
        int grid = (B + tpb - 1) / tpb;
        float t = time_ms([&] { kern<<<grid, tpb>>>(B, T, M, pi, Bm, obs, psi, dout); });
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hp.data(), psi, hp.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(hd.data(), dout, hd.size() * 8, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (size_t k = 0; k < hp.size(); k++) bad += hp[k] != hr[k];
        size_t badd = 0;
        for (size_t k = 0; k < hd.size(); k++) badd += hd[k] != hdr[k];
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, tpb, 0);
        printf("%-28s %.2f ms  %.3e tr/s  psi_mismatch=%zu delta_mismatch=%zu occ=%d %s\n", name, t, trans / t * 1e3,
               bad, badd, occ, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    auto runw = [&](const char *name, auto kern, int S, int W, int ctas_per_sm) {
        cudaMemset(psi, 0, (size_t)B * T * N);
        int grid = sms * ctas_per_sm;
        if (grid * S > B) grid = (B + S - 1) / S;
        float t = time_ms([&] { kern<<<grid, 32 * W>>>(B, T, M, pi, A, Bm, obs, psi, dout); });
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hp.data(), psi, hp.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(hd.data(), dout, hd.size() * 8, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (size_t k = 0; k < hp.size(); k++) bad += hp[k] != hr[k];
        size_t badd = 0;
        for (size_t k = 0; k < hd.size(); k++) badd += hd[k] != hdr[k];
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * W, 0);
        printf("%-28s %.2f ms  %.3e tr/s  psi_mismatch=%zu delta_mismatch=%zu occ=%d %s\n", name, t, trans / t * 1e3,
               bad, badd, occ, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    runw("TW S4 W4 CH1", vit64_tw_kernel<4, 4, 1>, 4, 4, 4);
    runw("TW S4 W4 CH2", vit64_tw_kernel<4, 4, 2>, 4, 4, 4);
    runw("TW S2 W4 CH2", vit64_tw_kernel<2, 4, 2>, 2, 4, 4);
    runw("TW S8 W4 CH1", vit64_tw_kernel<8, 4, 1>, 8, 4, 3);
    runw("TW S4 W2 CH1", vit64_tw_kernel<4, 2, 1>, 4, 2, 4);
    runw("TW S4 W4 CH1 x3", vit64_tw_kernel<4, 4, 1>, 4, 4, 3);
    run("T2 S4 CH1", vit64_t2_kernel<4, 1>, 4, 4);
    run("T2 S4 CH2", vit64_t2_kernel<4, 2>, 4, 4);
    run("T2 S2 CH2", vit64_t2_kernel<2, 2>, 2, 4);
    run("T2 S8 CH2", vit64_t2_kernel<8, 2>, 8, 4);
    run("T2 S4 CH4", vit64_t2_kernel<4, 4>, 4, 4);
    run("S4 V3 C2", vit64_kernel<4, 3, 2>, 4, 4);
    run("S4 V3 C4", vit64_kernel<4, 3, 4>, 4, 4);
    run("S4 V6 C2", vit64_kernel<4, 6, 2>, 4, 4);
    run("S4 V4 C2 (max only)", vit64_kernel<4, 4, 2>, 4, 4);
    run("S4 V5 C2 (adds only)", vit64_kernel<4, 5, 2>, 4, 4);
    run("S2 V3 C2", vit64_kernel<2, 3, 2>, 2, 4);
    return 0;
}
