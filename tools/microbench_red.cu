// How fast can the emission-count scatter go?  (Development tool.)  Every (sequence, step) adds a row
// of N=128 floats into row o_t of an L2-resident [M=1024][128] table.  Variants:
//   0: thread per row, red.global.add.v4.f32 (32 per row, lanes hit 32 different rows)
//   1: warp per row, lane = 4 consecutive columns (one coalesced 512-byte red.v4 per row)
//   2: warp per row, scalar red.f32 (4 coalesced instructions per row)
//   3: thread per row, scalar red.f32 (128 per row)
//   4: warp per row, red.v2.f32
//   5: like 1 but into R=8 replicas of the table chosen by blockIdx
//   bulk: TMA bulk reductions (cp.reduce.async.bulk) of whole rows staged in shared memory, into 8 replicas
// Prints rows/us and float adds per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_red tools/microbench_red.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ void red_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_v2(float *p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_1(float *p, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ uint32_t hash(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int MODE>
__global__ void __launch_bounds__(128) k(float *table, int rows_per_thread_or_warp, int M) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = tid >> 5;
    float v = 1e-6f * (lane + 1);
    if (MODE == 5) table += (size_t)(blockIdx.x & 7) * M * 128;
    for (int it = 0; it < rows_per_thread_or_warp; it++) {
        if (MODE == 0 || MODE == 3) {
            int row = hash(tid * 7919u + it) % M;
            float *p = table + (size_t)row * 128;
            if (MODE == 0) {
#pragma unroll
                for (int c = 0; c < 32; c++) red_v4(p + 4 * c, v, v, v, v);
            } else {
#pragma unroll
                for (int c = 0; c < 128; c++) red_1(p + c, v);
            }
        } else {
            int row = hash(warp * 7919u + it) % M;
            float *p = table + (size_t)row * 128;
            if (MODE == 1 || MODE == 5) red_v4(p + 4 * lane, v, v, v, v);
            else if (MODE == 4) { red_v2(p + 2 * lane, v, v); red_v2(p + 64 + 2 * lane, v, v); }
            else {
#pragma unroll
                for (int c = 0; c < 4; c++) red_1(p + 32 * c + lane, v);
            }
        }
    }
}

// TMA bulk reduction: rows staged in shared memory, `issuers` lanes of warp 0 each send whole 512-byte rows with
// cp.reduce.async.bulk.global.shared::cta.add.f32 (one instruction per row), groups of `per_group` rows per commit.
__global__ void __launch_bounds__(128) k_bulk(float *table, int rows_per_cta, int M, int issuers, int row_bytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    float *rows = reinterpret_cast<float *>(sm);            // 64 rows x 512 B staging
    for (int k = threadIdx.x; k < 64 * 128; k += blockDim.x) rows[k] = 1e-6f * (k & 31);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    table += (size_t)(blockIdx.x & 7) * M * 128;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32 && lane < issuers) {
        const int pieces = 512 / row_bytes;
        for (int it = lane; it < rows_per_cta; it += issuers) {
            int row = hash(blockIdx.x * 7919u + it) % M;
            for (int pc = 0; pc < pieces; pc++) {
                float *dst = table + (size_t)row * 128 + pc * (row_bytes / 4);
                uint32_t src = (uint32_t)__cvta_generic_to_shared(rows + (it & 63) * 128 + pc * (row_bytes / 4));
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                             ::"l"(dst), "r"(src), "r"(row_bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (((it / issuers) & 15) == 15) asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

static void run_bulk(const char *name, float *table, int M, int sms, double mhz, int issuers, int row_bytes) {
    const int ctas = sms, per = 16384;
    double rows = (double)ctas * per;
    k_bulk<<<ctas, 128, 64 * 512>>>(table, per / 8, M, issuers, row_bytes);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_bulk<<<ctas, 128, 64 * 512>>>(table, per, M, issuers, row_bytes);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.3f ms  %8.1f rows/us  %6.2f float-adds/clk/SM (at %.0f MHz)  err=%s\n", name, ms, rows / ms * 1e-3,
           rows * 128 / (ms * 1e-3) / sms / (mhz * 1e6), mhz, cudaGetErrorString(cudaGetLastError()));
}

template <int MODE>
static void run(const char *name, float *table, int M, int sms, double mhz) {
    const int ctas = sms * 8;
    const bool per_thread = MODE == 0 || MODE == 3;
    const int per = per_thread ? 64 : 2048;
    double rows = (double)ctas * (per_thread ? 128 : 4) * per;
    k<MODE><<<ctas, 128>>>(table, per, M);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<ctas, 128>>>(table, per, M);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.3f ms  %8.1f rows/us  %6.2f float-adds/clk/SM (at %.0f MHz)  err=%s\n", name, ms, rows / ms * 1e-3,
           rows * 128 / (ms * 1e-3) / sms / (mhz * 1e6), mhz, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int M = 1024;
    float *table;
    cudaMalloc(&table, (size_t)8 * M * 128 * 4);
    cudaMemset(table, 0, (size_t)8 * M * 128 * 4);
    double mhz = clk * 1e-3;
    run<0>("thread/row red.v4 (32 per row)", table, M, p.multiProcessorCount, mhz);
    run<1>("warp/row red.v4 coalesced", table, M, p.multiProcessorCount, mhz);
    run<4>("warp/row red.v2 coalesced", table, M, p.multiProcessorCount, mhz);
    run<2>("warp/row red.f32 coalesced", table, M, p.multiProcessorCount, mhz);
    run<3>("thread/row red.f32 (128 per row)", table, M, p.multiProcessorCount, mhz);
    run<5>("warp/row red.v4 coalesced, 8 replicas", table, M, p.multiProcessorCount, mhz);
    run_bulk("TMA bulk reduce 512 B rows, 1 issuer/CTA", table, M, p.multiProcessorCount, mhz, 1, 512);
    run_bulk("TMA bulk reduce 512 B rows, 4 issuers/CTA", table, M, p.multiProcessorCount, mhz, 4, 512);
    run_bulk("TMA bulk reduce 512 B rows, 32 issuers/CTA", table, M, p.multiProcessorCount, mhz, 32, 512);
    run_bulk("TMA bulk reduce 4 x 128 B pieces, 4 issuers", table, M, p.multiProcessorCount, mhz, 4, 128);
    run_bulk("TMA bulk reduce 4 x 128 B pieces, 32 issuers", table, M, p.multiProcessorCount, mhz, 32, 128);
    return 0;
}
