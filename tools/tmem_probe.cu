// Bring-up probes for the fused Baum-Welch tensor-core kernels (development tool):
//   1. register <-> TMEM mapping of tcgen05.ld.16x256b (used to turn thread-per-row data into
//      sector-coalesced red.global.add.v2.f32 rows);
//   2. tcgen05.mma kind::tf32 with the A operand in TMEM (written by tcgen05.st) and a K-major
//      128B-swizzled B operand written into shared memory BY THREADS (manual Swizzle<3,4,3>);
//   3. MN-major x MN-major MMA with both operands written by threads (manual Swizzle<2,5,2>);
//   4. TMEM load / store throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tmem_probe tools/tmem_probe.cu -lcuda
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../kerehmm_b200/csrc/tcgen05.cuh"

using namespace khmm;
using namespace khmm::tc;

struct __align__(8) Bars { uint64_t done; uint32_t tmem_base; };

// out_map[thread][reg] = value found (row*1000 + col) for the 16x256b.x4 load at lane offset 0 and 16
__global__ void __launch_bounds__(128, 1) probe_kernel(float *out_map, const float *Ah, const float *Bh, float *D1,
                                                       const float *Ph, const float *Qh, float *D2, long long *clk) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *tiles = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *sB = tiles;               // 64 KB K-major B operand / MN-major P
    unsigned char *sQ = tiles + 65536;       // 64 KB MN-major Q
    Bars *bars = (Bars *)(tiles + 131072);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
    if (threadIdx.x == 0) { mbar_init(&bars->done, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

    // ---- 1. mapping probe: region cols [384, 416) holds row*1000+col
    {
        uint32_t v[32];
#pragma unroll
        for (int q = 0; q < 32; q++) v[q] = __float_as_uint((float)(row * 1000 + q));
        tmem_st_32x32(tmem + lane_base + 384, v);
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        uint32_t r[16];
        for (int half = 0; half < 2; half++) {
            tmem_ld_16x256_x4(tmem + lane_base + ((uint32_t)(half * 16) << 16) + 384, r);
            tmem_ld_wait();
            for (int q = 0; q < 16; q++) out_map[(threadIdx.x * 2 + half) * 16 + q] = __uint_as_float(r[q]);
        }
    }
    // ---- 2. A (TMEM, cols [128,256)) x B (smem K-major SW128 written by threads) -> D cols [0,128)
    {
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
#pragma unroll
            for (int q = 0; q < 32; q++) v[q] = __float_as_uint(Ah[row * 128 + c0 + q]);
            tmem_st_32x32(tmem + lane_base + 128 + c0, v);
        }
        tmem_st_wait();
        // B[n][k]: thread n writes its row
        for (int k = 0; k < 128; k += 4) {
            int kb = k >> 5, ch = (k & 31) >> 2;
            float4 val = *reinterpret_cast<const float4 *>(Bh + row * 128 + k);
            *reinterpret_cast<float4 *>(sB + kb * 16384 + row * 128 + ((ch ^ (row & 7)) << 4)) = val;
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (threadIdx.x == 0) {
            const uint32_t idesc = make_idesc_tf32(128, 128);
            for (int k = 0; k < 16; k++) {
                uint64_t db = make_kmajor_sw128_desc(sB + (k >> 2) * 16384);
                mma_tf32_ts(tmem, tmem + 128 + k * 8, desc_advance(db, (k & 3) * 32), idesc, k != 0);
            }
            tc_commit(&bars->done);
        }
        mbar_wait(&bars->done, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem + lane_base + c0, r);
            tmem_ld_wait();
            for (int q = 0; q < 32; q++) D1[row * 128 + c0 + q] = __uint_as_float(r[q]);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ---- 3. MN-major P (sB) x MN-major Q (sQ), both written by threads with Swizzle<2,5,2>
    {
        for (int c = 0; c < 128; c += 4) {
            int sub = c >> 5, c32 = (c & 31) >> 3, h = (c & 7) >> 2;
            int off = sub * 16384 + row * 128 + (((c32 ^ (row & 3)) << 5) | (h << 4));
            *reinterpret_cast<float4 *>(sB + off) = *reinterpret_cast<const float4 *>(Ph + row * 128 + c);
            *reinterpret_cast<float4 *>(sQ + off) = *reinterpret_cast<const float4 *>(Qh + row * 128 + c);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (threadIdx.x == 0) {
            const uint32_t idesc = make_idesc_tf32(128, 128, true, true);
            uint64_t da = make_mnmajor_sw128a32_desc(sB, 16384, 512);
            uint64_t db = make_mnmajor_sw128a32_desc(sQ, 16384, 512);
            for (int k = 0; k < 16; k++)
                mma_tf32_ss(tmem + 256, desc_advance(da, k * 1024), desc_advance(db, k * 1024), idesc, k != 0);
            tc_commit(&bars->done);
        }
        mbar_wait(&bars->done, 1);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem + lane_base + 256 + c0, r);
            tmem_ld_wait();
            for (int q = 0; q < 32; q++) D2[row * 128 + c0 + q] = __uint_as_float(r[q]);
        }
    }
    // ---- 4. TMEM ld / st throughput (all 4 warps, 128 columns each iteration)
    {
        __syncthreads();
        long long t0 = clock64();
        uint32_t acc = 0;
        for (int it = 0; it < 64; it++) {
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem + lane_base + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 32; q++) acc ^= r[q];
            }
        }
        __syncthreads();
        long long t1 = clock64();
        for (int it = 0; it < 64; it++) {
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t v[32];
#pragma unroll
                for (int q = 0; q < 32; q++) v[q] = acc + q + it;
                tmem_st_32x32(tmem + lane_base + 384 + (c0 & 96), v);
            }
            tmem_st_wait();
        }
        __syncthreads();
        long long t2 = clock64();
        if (threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = acc; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static float tf32_trunc(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

int main() {
    std::vector<float> hA(128 * 128), hB(128 * 128), hP(128 * 128), hQ(128 * 128), hD1(128 * 128), hD2(128 * 128),
        hmap(128 * 2 * 16);
    srand(3);
    for (auto &v : hA) v = tf32_trunc(rand() / (float)RAND_MAX - 0.3f);
    for (auto &v : hB) v = tf32_trunc(rand() / (float)RAND_MAX - 0.5f);
    for (auto &v : hP) v = tf32_trunc(rand() / (float)RAND_MAX - 0.4f);
    for (auto &v : hQ) v = tf32_trunc(rand() / (float)RAND_MAX - 0.6f);
    float *dA, *dB, *dP, *dQ, *dD1, *dD2, *dmap;
    long long *dclk;
    size_t sz = 128 * 128 * 4;
    cudaMalloc(&dA, sz); cudaMalloc(&dB, sz); cudaMalloc(&dP, sz); cudaMalloc(&dQ, sz); cudaMalloc(&dD1, sz);
    cudaMalloc(&dD2, sz); cudaMalloc(&dmap, hmap.size() * 4); cudaMalloc(&dclk, 64);
    cudaMemcpy(dA, hA.data(), sz, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), sz, cudaMemcpyHostToDevice);
    cudaMemcpy(dP, hP.data(), sz, cudaMemcpyHostToDevice);
    cudaMemcpy(dQ, hQ.data(), sz, cudaMemcpyHostToDevice);
    size_t smem = 131072 + sizeof(Bars) + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 128, smem>>>(dmap, dA, dB, dD1, dP, dQ, dD2, dclk);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    long long hclk[3];
    cudaMemcpy(hmap.data(), dmap, hmap.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hD1.data(), dD1, sz, cudaMemcpyDeviceToHost);
    cudaMemcpy(hD2.data(), dD2, sz, cudaMemcpyDeviceToHost);
    cudaMemcpy(hclk, dclk, 24, cudaMemcpyDeviceToHost);
    printf("16x256b.x4 mapping (value = row*1000+col), threads 0..7 and 32..35, lane-offset 0 then 16:\n");
    for (int t : {0, 1, 2, 3, 4, 5, 8, 31, 32, 33}) {
        for (int half = 0; half < 2; half++) {
            printf("  t%-3d h%d:", t, half);
            for (int q = 0; q < 16; q++) printf(" %6.0f", hmap[(t * 2 + half) * 16 + q]);
            printf("\n");
        }
    }
    // check the conjectured mapping: reg[4n+e] -> row (l/4), col 8n + 2(l%4) + e ; reg[4n+2+e] -> row l/4+8
    int bad = 0;
    for (int t = 0; t < 128; t++)
        for (int half = 0; half < 2; half++)
            for (int n = 0; n < 4; n++)
                for (int e = 0; e < 4; e++) {
                    int l = t & 31, w = t >> 5;
                    int row = w * 32 + half * 16 + (l >> 2) + ((e >> 1) ? 8 : 0);
                    int col = 8 * n + 2 * (l & 3) + (e & 1);
                    if (hmap[(t * 2 + half) * 16 + 4 * n + e] != (float)(row * 1000 + col)) bad++;
                }
    printf("conjectured 16x256b mapping mismatches: %d\n", bad);
    double e1 = 0, e2 = 0, m1 = 0, m2 = 0;
    for (int i = 0; i < 128; i++)
        for (int j = 0; j < 128; j++) {
            double a = 0, b = 0;
            for (int k = 0; k < 128; k++) {
                a += (double)hA[i * 128 + k] * hB[j * 128 + k];
                b += (double)hP[k * 128 + i] * hQ[k * 128 + j];
            }
            e1 = fmax(e1, fabs(a - hD1[i * 128 + j])); m1 = fmax(m1, fabs(a));
            e2 = fmax(e2, fabs(b - hD2[i * 128 + j])); m2 = fmax(m2, fabs(b));
        }
    printf("TMEM-A x smem-B(K-major, thread-written): max|ref|=%.3f max err=%.3e\n", m1, e1);
    printf("MN x MN (thread-written Swizzle<2,5,2>):   max|ref|=%.3f max err=%.3e\n", m2, e2);
    printf("TMEM ld: %.1f clk per 128x128 fp32 tile (4 warps) ; st: %.1f clk per tile\n", hclk[0] / 64.0, hclk[1] / 64.0);
    return (bad == 0 && e1 < 1e-3 * m1 && e2 < 1e-3 * m2) ? 0 : 2;
}
