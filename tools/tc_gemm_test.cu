// Bring-up test for the tcgen05 tf32 GEMM building block (development tool):
//   D[M][N] = X[M][K] * W[N][K]^T, fp32 in, tf32 multiply, fp32 accumulate in TMEM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc_gemm_test tools/tc_gemm_test.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../kerehmm_b200/csrc/tcgen05.cuh"

using namespace khmm;
using namespace khmm::tc;

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 4;
constexpr int TILE_BYTES = BM * BK * 4;  // 16 KB per operand per stage

struct __align__(8) Bars {
    uint64_t full[STAGES], empty[STAGES], tmem_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(192, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, float *__restrict__ D,
                 int M, int N, int K) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    Bars *bars = (Bars *)(tiles + STAGES * 2 * TILE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int nkb = K / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(&bars->tmem_base, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(&mapX);
            tma_prefetch_desc(&mapW);
            for (int kb = 0; kb < nkb; kb++) {
                int s = kb % STAGES;
                uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&bars->empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&bars->full[s], 2 * TILE_BYTES);
                tma_load_2d(tiles + s * 2 * TILE_BYTES, &mapX, &bars->full[s], kb * BK, m0);
                tma_load_2d(tiles + s * 2 * TILE_BYTES + TILE_BYTES, &mapW, &bars->full[s], kb * BK, n0);
            }
        }
    } else if (warp == 5) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_tf32(BM, BN);
            for (int kb = 0; kb < nkb; kb++) {
                int s = kb % STAGES;
                uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&bars->full[s], ph);
                tc_fence_after();
                uint64_t da = make_kmajor_sw128_desc(tiles + s * 2 * TILE_BYTES);
                uint64_t db = make_kmajor_sw128_desc(tiles + s * 2 * TILE_BYTES + TILE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 8; k++)
                    mma_tf32_ss(tmem, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc, (kb | k) != 0);
                tc_commit(&bars->empty[s]);
            }
            tc_commit(&bars->tmem_full);
        }
    } else {
        mbar_wait(&bars->tmem_full, 0);
        tc_fence_after();
        const int row = m0 + warp * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
            tmem_ld_wait();
            if (row < M) {
                float4 *dst = reinterpret_cast<float4 *>(D + (size_t)row * N + n0 + c);
#pragma unroll
                for (int q = 0; q < 8; q++)
                    dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                         __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, 128);
}

// ---- MN-major x MN-major: D[M][N] = sum_k P[k][m] * Q[k][n]; P is [K][M] row-major, Q is [K][N] row-major.
// A k-block is 128 rows (K) x 128 columns, staged by TMA as 4 sub-tiles of [128 k-rows][32 cols] (128 B rows,
// SWIZZLE_128B).  As an MN-major UMMA operand one sub-tile is: 32 MN elements contiguous (128 B), 8 k-rows per
// 1024-byte swizzle atom, atoms of the next 8 k-rows 1024 B further (SBO), the next 32 MN elements in the next
// sub-tile, 16 KB further (LBO).
constexpr int XS = 3;        // stages
constexpr int XR = 64;       // K rows per k-block
constexpr int XSUB = XR * 128;          // bytes of one [XR][32 fp32] sub-tile
constexpr int XOP = 4 * XSUB;           // one operand k-block: 4 sub-tiles (128 MN columns)
struct __align__(8) XBars { uint64_t full[XS], empty[XS], tmem_full; uint32_t tmem_base; };

__global__ void __launch_bounds__(192, 1)
gemm_mn_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapQ, float *__restrict__ D,
               int K, int ldd, int q_row_shift) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    XBars *bars = (XBars *)(tiles + XS * 2 * XOP);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = K / XR;
    if (threadIdx.x == 0) {
        for (int s = 0; s < XS; s++) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(&bars->tmem_base, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    if (warp == 4) {
        if (elect_one()) {
            for (int kb = 0; kb < nkb; kb++) {
                int s = kb % XS;
                uint32_t ph = (kb / XS) & 1;
                mbar_wait(&bars->empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&bars->full[s], 2 * XOP);
                for (int c = 0; c < 4; c++) {
                    tma_load_2d(tiles + s * 2 * XOP + c * XSUB, &mapP, &bars->full[s], c * 32, kb * XR);
                    tma_load_2d(tiles + s * 2 * XOP + XOP + c * XSUB, &mapQ, &bars->full[s], c * 32, kb * XR + q_row_shift);
                }
            }
        }
    } else if (warp == 5) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_tf32(128, 128, true, true);
            for (int kb = 0; kb < nkb; kb++) {
                int s = kb % XS;
                uint32_t ph = (kb / XS) & 1;
                mbar_wait(&bars->full[s], ph);
                tc_fence_after();
                uint64_t da = make_mnmajor_sw128a32_desc(tiles + s * 2 * XOP, XSUB, 512);
                uint64_t db = make_mnmajor_sw128a32_desc(tiles + s * 2 * XOP + XOP, XSUB, 512);
#pragma unroll
                for (int k = 0; k < XR / 8; k++)   // 8 k-rows (two 512-byte atoms) per MMA
                    mma_tf32_ss(tmem, desc_advance(da, k * 1024), desc_advance(db, k * 1024), idesc, (kb | k) != 0);
                tc_commit(&bars->empty[s]);
            }
            tc_commit(&bars->tmem_full);
        }
    } else {
        mbar_wait(&bars->tmem_full, 0);
        tc_fence_after();
        const int row = warp * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 32; q++) D[(size_t)row * ldd + c + q] = __uint_as_float(r[q]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, 128);
}

static float tf32_trunc(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

int main(int argc, char **argv) {
    int M = argc > 1 ? atoi(argv[1]) : 256, N = argc > 2 ? atoi(argv[2]) : 256, K = argc > 3 ? atoi(argv[3]) : 128;
    std::vector<float> hX((size_t)M * K), hW((size_t)N * K), hD((size_t)M * N);
    srand(1);
    for (auto &v : hX) v = (rand() / (float)RAND_MAX) - 0.3f;
    for (auto &v : hW) v = (rand() / (float)RAND_MAX) - 0.5f;
    float *X, *W, *D;
    cudaMalloc(&X, hX.size() * 4); cudaMalloc(&W, hW.size() * 4); cudaMalloc(&D, hD.size() * 4);
    cudaMemcpy(X, hX.data(), hX.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(W, hW.data(), hW.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(D, 0xff, hD.size() * 4);
    CUtensorMap mX, mW;
    int r1 = make_tmap_f32_rowmajor(&mX, X, M, K, BM), r2 = make_tmap_f32_rowmajor(&mW, W, N, K, BN);
    if (r1 || r2) { printf("tensor map creation failed %d %d\n", r1, r2); return 1; }
    size_t smem = STAGES * 2 * TILE_BYTES + sizeof(Bars) + 1024;
    cudaFuncSetAttribute(gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(M / BM, N / BN);
    gemm_tf32_kernel<<<grid, 192, smem>>>(mX, mW, D, M, N, K);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(hD.data(), D, hD.size() * 4, cudaMemcpyDeviceToHost);
    double max_err_tf32 = 0, max_err_fp32 = 0, max_ref = 0;
    for (int i = 0; i < M; i += 7)
        for (int j = 0; j < N; j += 5) {
            double a = 0, b = 0;
            for (int k = 0; k < K; k++) {
                a += (double)tf32_trunc(hX[(size_t)i * K + k]) * tf32_trunc(hW[(size_t)j * K + k]);
                b += (double)hX[(size_t)i * K + k] * hW[(size_t)j * K + k];
            }
            double d = hD[(size_t)i * N + j];
            max_err_tf32 = fmax(max_err_tf32, fabs(d - a));
            max_err_fp32 = fmax(max_err_fp32, fabs(d - b));
            max_ref = fmax(max_ref, fabs(b));
        }
    printf("M=%d N=%d K=%d  max|ref|=%.3f  max err vs tf32-truncated ref=%.3e  vs fp32 ref=%.3e\n", M, N, K, max_ref,
           max_err_tf32, max_err_fp32);
    {   // MN-major test: P [K2][128], Q [K2][128] -> D2[128][128]
        int K2 = 1024;
        std::vector<float> hP((size_t)(K2 + 1) * 128), hQ((size_t)(K2 + 1) * 128), hD2(128 * 128);
        for (auto &v : hP) v = tf32_trunc((rand() / (float)RAND_MAX) - 0.3f);
        for (auto &v : hQ) v = tf32_trunc((rand() / (float)RAND_MAX) - 0.5f);
        float *P, *Q, *D2;
        cudaMalloc(&P, hP.size() * 4); cudaMalloc(&Q, hQ.size() * 4); cudaMalloc(&D2, hD2.size() * 4);
        cudaMemcpy(P, hP.data(), hP.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(Q, hQ.data(), hQ.size() * 4, cudaMemcpyHostToDevice);
        CUtensorMap mP, mQ;
        int rp = make_tmap_f32_rowmajor(&mP, P, K2 + 1, 128, XR, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        int rq = make_tmap_f32_rowmajor(&mQ, Q, K2 + 1, 128, XR, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        printf("mn-major tensor maps: %d %d\n", rp, rq);
        cudaMemset(D2, 0xff, hD2.size() * 4);
        size_t sm2 = XS * 2 * XOP + sizeof(XBars) + 1024;
        cudaFuncSetAttribute(gemm_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
        gemm_mn_kernel<<<1, 192, sm2>>>(mP, mQ, D2, K2, 128, 1);
        cudaError_t e2 = cudaGetLastError();
        if (e2 == cudaSuccess) e2 = cudaDeviceSynchronize();
        printf("mn-major launch: %s\n", cudaGetErrorString(e2));
        if (e2 != cudaSuccess) return 1;
        cudaMemcpy(hD2.data(), D2, hD2.size() * 4, cudaMemcpyDeviceToHost);
        double maxe = 0, maxr = 0;
        for (int i = 0; i < 128; i++)
            for (int j = 0; j < 128; j++) {
                double a = 0;
                for (int k = 0; k < K2; k++) a += (double)hP[(size_t)k * 128 + i] * hQ[(size_t)(k + 1) * 128 + j];
                maxe = fmax(maxe, fabs(hD2[i * 128 + j] - a));
                maxr = fmax(maxr, fabs(a));
            }
        printf("mn-major (Q shifted by one row): max|ref|=%.3f max err=%.3e\n", maxr, maxe);
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 6; j++) {
                double a = 0;
                for (int k = 0; k < K2; k++) a += (double)hP[(size_t)k * 128 + i] * hQ[(size_t)(k + 1) * 128 + j];
                printf("  D[%d][%d]=%.4f ref=%.4f", i, j, hD2[i * 128 + j], a);
            }
            printf("\n");
        }
        if (!(maxe < 1e-3 * maxr)) return 3;
    }
    // timing
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int it = 0; it < 20; it++) gemm_tf32_kernel<<<grid, 192, smem>>>(mX, mW, D, M, N, K);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%.3f us per GEMM, %.1f TFLOP/s\n", ms / 20 * 1e3, 2.0 * M * N * K / (ms / 20 * 1e-3) * 1e-12);
    return max_err_tf32 < 1e-2 * max_ref ? 0 : 2;
}
