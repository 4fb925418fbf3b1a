// tcgen05.mma issue-rate microbenchmark (development tool): the tensor-pipe peaks that the c4 / c5 rooflines are
// quoted against.  One CTA per SM; operands sit in shared memory (K-major, 128-byte swizzle; contents irrelevant),
// one elected thread issues `rounds` x (K = 128) MMAs into one TMEM accumulator and commits once.
//   kinds: tf32 SS (both operands in shared memory), tf32 TS (A operand in TMEM, as the c4 kernels use it),
//          bf16 SS (kind::f16); tile 128 x 128 and 128 x 256.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench_mma tools/microbench_mma.cu
//   tools/microbench_mma > profiles/r2_microbench_tf32.json
#include <stdio.h>
#include <stdlib.h>

#include "../kerehmm_b200/csrc/tcgen05.cuh"

using namespace khmm;
using namespace khmm::tc;

struct __align__(8) Bars {
    uint64_t done;
    uint32_t tmem_base;
};

// mode 0: tf32 SS, 1: tf32 TS, 2: bf16 SS
template <int MODE, int BN>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int rounds, float *sink) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sA = tiles;                     // 128 rows x 128 K fp32 (64 KB) or bf16 (32 KB)
    unsigned char *sB = tiles + 64 * 1024;         // BN rows x 128 K
    Bars *bars = (Bars *)(sB + (size_t)BN * 128 * 4);
    const int warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < (64 * 1024 + BN * 128 * 4) / 4; k += blockDim.x) ((uint32_t *)tiles)[k] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(&bars->done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&bars->tmem_base, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = MODE == 2 ? make_idesc_bf16(128, BN) : make_idesc_tf32(128, BN);
        for (int r = 0; r < rounds; r++) {
            if (MODE == 2) {
#pragma unroll
                for (int k = 0; k < 8; k++) {      // K = 16 bf16 per instruction, 2 sub-tiles of 64 elements
                    uint64_t da = desc_advance(make_kmajor_sw128_desc(sA + (k >> 2) * 128 * 128), (k & 3) * 32);
                    uint64_t db = desc_advance(make_kmajor_sw128_desc(sB + (k >> 2) * BN * 128), (k & 3) * 32);
                    mma_bf16_ss(tmem, da, db, idesc, true);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++) {     // K = 8 tf32 per instruction, 4 sub-tiles of 32 elements
                    uint64_t db = desc_advance(make_kmajor_sw128_desc(sB + (k >> 2) * BN * 128), (k & 3) * 32);
                    if (MODE == 1) {
                        mma_tf32_ts(tmem, tmem + 256 + k * 8, db, idesc, true);
                    } else {
                        uint64_t da = desc_advance(make_kmajor_sw128_desc(sA + (k >> 2) * 128 * 128), (k & 3) * 32);
                        mma_tf32_ss(tmem, da, db, idesc, true);
                    }
                }
            }
        }
        tc_commit(&bars->done);
        mbar_wait(&bars->done, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && sink == (float *)1) sink[0] = 0.f;
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int MODE, int BN>
static double run(int sms, int rounds) {
    size_t smem = 64 * 1024 + (size_t)BN * 128 * 4 + sizeof(Bars) + 1024;
    cudaFuncSetAttribute(mma_rate_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    mma_rate_kernel<MODE, BN><<<sms, 128, smem>>>(rounds / 8, nullptr);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int it = 0; it < 5; it++) {
        cudaEventRecord(e0);
        mma_rate_kernel<MODE, BN><<<sms, 128, smem>>>(rounds, nullptr);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "mode %d BN %d: %s\n", MODE, BN, cudaGetErrorString(e));
        return -1.0;
    }
    double flop = (double)sms * rounds * 2.0 * 128.0 * BN * 128.0;
    return flop / (best * 1e-3) / 1e12;
}

int main(int argc, char **argv) {
    int rounds = argc > 1 ? atoi(argv[1]) : 20000;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double tf_ss128 = run<0, 128>(sms, rounds), tf_ss256 = run<0, 256>(sms, rounds);
    double tf_ts128 = run<1, 128>(sms, rounds), tf_ts256 = run<1, 256>(sms, rounds);
    double bf_ss128 = run<2, 128>(sms, rounds), bf_ss256 = run<2, 256>(sms, rounds);
    double best_tf = tf_ss128;
    if (tf_ss256 > best_tf) best_tf = tf_ss256;
    if (tf_ts128 > best_tf) best_tf = tf_ts128;
    if (tf_ts256 > best_tf) best_tf = tf_ts256;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"rounds\": %d, \"how\": \"one CTA per SM issuing back-to-back tcgen05.mma "
           "(cta_group::1, M=128, K=128 per round) on resident shared-memory/TMEM operands, best of 5, CUDA events\",\n"
           " \"tf32_tflops\": %.1f, \"tf32_ss_n128\": %.1f, \"tf32_ss_n256\": %.1f, \"tf32_ts_n128\": %.1f, "
           "\"tf32_ts_n256\": %.1f,\n \"bf16_ss_n128\": %.1f, \"bf16_ss_n256\": %.1f}\n",
           p.name, sms, rounds, best_tf, tf_ss128, tf_ss256, tf_ts128, tf_ts256, bf_ss128, bf_ss256);
    return 0;
}
