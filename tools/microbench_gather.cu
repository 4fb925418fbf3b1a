// Probe (development tool): row gathers from an L2-resident [M][128] fp32 table into shared memory through the TMA
// unit instead of the LSU -- the emission-row gather of the tensor-core E-step.  Variants:
//   gather4:  cp.async.bulk.tensor.2d ... tile::gather4 (4 table rows per instruction), tensor map box {128, boxrows}
//   bulk:     cp.async.bulk.shared.global (one 512-byte row per instruction), issued by `issuers` lanes
// Checks the data and prints rows/us per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench_gather tools/microbench_gather.cu
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../kerehmm_b200/csrc/tcgen05.cuh"

using namespace khmm;
using namespace khmm::tc;

__device__ __forceinline__ uint32_t hash(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__device__ __forceinline__ void tma_gather4(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int r0, int r1,
                                            int r2, int r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mode 0: gather4 by one thread; mode 1: bulk rows by `issuers` lanes.  Each round fills a slab of 32 rows x 512 B.
__global__ void __launch_bounds__(128) k_gather(const __grid_constant__ CUtensorMap map, const float *table, int M, int rounds,
                                                int mode, int issuers, int *errors) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float *slab = reinterpret_cast<float *>(sm);                 // 2 slabs x [32][128]
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 2 * 16384);
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int bad = 0;
    if (threadIdx.x < 32) {
        auto issue = [&](int r) {
            const int s = r & 1;
            if (lane == 0) mbar_arrive_expect_tx(&bar[s], 16384);
            __syncwarp();
            if (mode == 0) {
                if (lane < 8) {
                    int rows[4];
                    for (int k = 0; k < 4; k++) rows[k] = hash(blockIdx.x * 7919u + r * 32 + lane * 4 + k) % M;
                    tma_gather4(slab + s * 4096 + lane * 4 * 128, &map, &bar[s], 0, rows[0], rows[1], rows[2], rows[3]);
                }
            } else if (lane < issuers) {
                for (int k = lane; k < 32; k += issuers) {
                    int row = hash(blockIdx.x * 7919u + r * 32 + k) % M;
                    bulk_load(slab + s * 4096 + k * 128, table + (size_t)row * 128, 512, &bar[s]);
                }
            }
        };
        issue(0);
        for (int r = 0; r < rounds; r++) {
            const int s = r & 1;
            if (r + 1 < rounds) issue(r + 1);        // one slab in flight while the previous one is consumed
            mbar_wait(&bar[s], (r >> 1) & 1);
            // check one element per lane: row `lane`, column lane
            int row = hash(blockIdx.x * 7919u + r * 32 + lane) % M;
            float got = slab[s * 4096 + lane * 128 + lane];
            if (got != (float)(row * 128 + lane)) bad++;
            __syncwarp();
        }
        if (bad) atomicAdd(errors, bad);
    }
}

int main() {
    const int M = 1024;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    std::vector<float> h((size_t)M * 128);
    for (size_t k = 0; k < h.size(); k++) h[k] = (float)k;
    float *table;
    int *errors;
    cudaMalloc(&table, h.size() * 4);
    cudaMalloc(&errors, 4);
    cudaMemcpy(table, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    PFN_encodeTiled encode = (PFN_encodeTiled)fn;
    const int rounds = 4000;
    size_t smem = 2 * 16384 + 64;
    cudaFuncSetAttribute(k_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int boxrows = 1; boxrows <= 4; boxrows += 3) {
        CUtensorMap map;
        cuuint64_t dims[2] = {128, (cuuint64_t)M};
        cuuint64_t strides[1] = {128 * 4};
        cuuint32_t box[2] = {128, (cuuint32_t)boxrows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            printf("gather4 box rows %d: tensor map creation failed (%d)\n", boxrows, (int)r);
            continue;
        }
        cudaMemset(errors, 0, 4);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_gather<<<p.multiProcessorCount, 128, smem>>>(map, table, M, 16, 0, 1, errors);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("gather4 box rows %d: kernel failed: %s\n", boxrows, cudaGetErrorString(e));
            return 1;
        }
        cudaMemset(errors, 0, 4);
        cudaEventRecord(e0);
        k_gather<<<p.multiProcessorCount, 128, smem>>>(map, table, M, rounds, 0, 1, errors);
        cudaEventRecord(e1);
        e = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        int herr = -1;
        cudaMemcpy(&herr, errors, 4, cudaMemcpyDeviceToHost);
        printf("gather4 (box rows %d), 8 instructions per 32-row slab: %.3f ms  %.1f rows/us/SM  %.1f cycles per slab  errors %d  (%s)\n",
               boxrows, ms, rounds * 32.0 / (ms * 1e3), ms * 1e-3 * 1.965e9 / rounds, herr, cudaGetErrorString(e));
    }
    CUtensorMap dummy;
    memset(&dummy, 0, sizeof(dummy));
    for (int issuers = 1; issuers <= 32; issuers *= 4) {
        if (issuers == 16) issuers = 32;
        cudaMemset(errors, 0, 4);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_gather<<<p.multiProcessorCount, 128, smem>>>(dummy, table, M, rounds, 1, issuers, errors);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        int herr = -1;
        cudaMemcpy(&herr, errors, 4, cudaMemcpyDeviceToHost);
        printf("bulk rows, %2d issuing lanes: %.3f ms  %.1f rows/us/SM  %.1f cycles per slab  errors %d  (%s)\n", issuers, ms,
               rounds * 32.0 / (ms * 1e3), ms * 1e-3 * 1.965e9 / rounds, herr, cudaGetErrorString(e));
    }
    return 0;
}
