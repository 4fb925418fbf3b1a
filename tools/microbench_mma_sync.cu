// Development: issue rate and latency of the warp-level mma.sync shapes on sm_100a (tf32 m16n8k8, bf16 m16n8k16),
// as used by the small-N log-likelihood kernel (smalln.cu).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb tools/microbench_mma_sync.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND, int CHAINS>
__global__ void k(int iters, float *out, long long *cycles) {
    float d[CHAINS][4];
    for (int c = 0; c < CHAINS; c++) for (int i = 0; i < 4; i++) d[c][i] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x3f800000u, b1 = 0x3f800000u;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(b0));
            else
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(b0));
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int c = 0; c < CHAINS; c++) for (int i = 0; i < 4; i++) s += d[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int KIND, int CHAINS>
void run(const char *name, int warps_per_sm_part) {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096, threads = 128 * warps_per_sm_part;   // 4 schedulers x warps
    k<KIND, CHAINS><<<148, threads>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    k<KIND, CHAINS><<<148, threads>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-22s chains %d warps/scheduler %d: %.1f cycles per mma per warp, %.1f cycles per mma per scheduler\n", name, CHAINS,
           warps_per_sm_part, (double)h / iters / CHAINS, (double)h / iters / CHAINS / warps_per_sm_part);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 1>("tf32 m16n8k8", 1);  run<0, 4>("tf32 m16n8k8", 1);  run<0, 4>("tf32 m16n8k8", 4);  run<0, 1>("tf32 m16n8k8", 8);
    run<3, 1>("tf32 m16n8k4", 1);  run<3, 4>("tf32 m16n8k4", 1);  run<3, 4>("tf32 m16n8k4", 4);
    run<1, 1>("bf16 m16n8k16", 1); run<1, 4>("bf16 m16n8k16", 1); run<1, 4>("bf16 m16n8k16", 4); run<1, 1>("bf16 m16n8k16", 8);
    run<2, 1>("bf16 m16n8k8", 1);  run<2, 4>("bf16 m16n8k8", 1);  run<2, 4>("bf16 m16n8k8", 4);
    return 0;
}
