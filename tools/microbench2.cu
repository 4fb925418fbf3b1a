// How fast can an 8x8 matrix-vector recursion issue?  Register- vs constant-bank-resident matrix,
// scalar FFMA vs packed FFMA2.  (Development tool.)  Prints FMA lane-ops per clock per SM.
#include <cuda_runtime.h>
#include <stdio.h>
#define ITERS 2048
__constant__ float cM[64];
__constant__ float2 cM2[32];

template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, const float *Mg, float seed) {
    float u[8];
#pragma unroll
    for (int i = 0; i < 8; i++) u[i] = seed + threadIdx.x * 1e-3f + i;
    float m[64];
    float2 m2[32];
    if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 64; i++) m[i] = Mg[i];
    }
    if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 32; i++) m2[i] = make_float2(Mg[2 * i], Mg[2 * i + 1]);
    }
    for (int it = 0; it < ITERS; it++) {
        float v[8];
        if (MODE == 0 || MODE == 1) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) s = fmaf(u[i], MODE == 0 ? m[i * 8 + j] : cM[i * 8 + j], s);
                v[j] = s;
            }
        } else {
            float2 o[4];
#pragma unroll
            for (int kq = 0; kq < 4; kq++) {
                float2 c0 = make_float2(0.f, 0.f), c1 = make_float2(0.f, 0.f);
#pragma unroll
                for (int y = 0; y < 8; y += 2) {
                    c0 = __ffma2_rn(make_float2(u[y], u[y]), MODE == 2 ? m2[y * 4 + kq] : cM2[y * 4 + kq], c0);
                    c1 = __ffma2_rn(make_float2(u[y + 1], u[y + 1]), MODE == 2 ? m2[(y + 1) * 4 + kq] : cM2[(y + 1) * 4 + kq], c1);
                }
                o[kq] = __fadd2_rn(c0, c1);
            }
#pragma unroll
            for (int kq = 0; kq < 4; kq++) { v[2 * kq] = o[kq].x; v[2 * kq + 1] = o[kq].y; }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) u[j] = v[j] * 0.124f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    float h[64];
    for (int i = 0; i < 64; i++) h[i] = 0.1f + 0.001f * i;
    float *Mg, *out;
    cudaMalloc(&Mg, 256); cudaMalloc(&out, 148 * 64 * 128 * 4);
    cudaMemcpy(Mg, h, 256, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(cM, h, 256);
    cudaMemcpyToSymbol(cM2, h, 256);
    const char *names[4] = {"scalar FFMA, matrix in registers", "scalar FFMA, matrix in constant bank",
                            "packed FFMA2, matrix in registers", "packed FFMA2, matrix in constant bank"};
    for (int wps = 1; wps <= 16; wps *= 2) {   // warps per SM sub-partition
        int grid = 148 * wps, block = 128;
        float t[4];
        t[0] = time_ms([&] { k<0><<<grid, block>>>(out, Mg, 1.f); });
        t[1] = time_ms([&] { k<1><<<grid, block>>>(out, Mg, 1.f); });
        t[2] = time_ms([&] { k<2><<<grid, block>>>(out, Mg, 1.f); });
        t[3] = time_ms([&] { k<3><<<grid, block>>>(out, Mg, 1.f); });
        for (int m = 0; m < 4; m++) {
            double fma = (double)grid * block * ITERS * 64;
            printf("warps/SMSP=%2d  %-40s %.3f ms  %.1f FMA/clk/SM (at 1.965 GHz)  cycles/warp-step=%.0f\n", wps, names[m], t[m],
                   fma / (t[m] * 1e-3) / 1.965e9 / 148, t[m] * 1e-3 * 1.965e9 / ITERS / 1.0);
        }
    }
    return 0;
}
