// Pipe-rate microbenchmarks for the roofline denominators SURVEY.md 8(d) says are not in
// MEASURED_PEAKS.json: FP32 FMA (scalar FFMA, packed FFMA2, and both mixed), MUFU ex2, FP64 add,
// FP64 add+compare+select (the Viterbi inner step).  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(256) k_fp32(float *out, float a, float b) {
    // 16 independent accumulators (scalar) or 8 packed pairs
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 av = make_float2(a, a * 1.0001f), bv = make_float2(b, b * 0.9999f);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) {  // 2 scalar FFMA
                acc[i].x = fmaf(acc[i].x, av.x, bv.x);
                acc[i].y = fmaf(acc[i].y, av.y, bv.y);
            } else if (MODE == 1) {  // 1 packed FFMA2
                acc[i] = __ffma2_rn(acc[i], av, bv);
            } else {  // mixed: even i packed, odd i scalar
                if (i & 1) {
                    acc[i].x = fmaf(acc[i].x, av.x, bv.x);
                    acc[i].y = fmaf(acc[i].y, av.y, bv.y);
                } else {
                    acc[i] = __ffma2_rn(acc[i], av, bv);
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_mufu(float *out, float a) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x * 1e-3f + i * 0.01f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float y;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc[i]));
            acc[i] = y - a;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_fp64(double *out, double a) {
    double acc[8];
    int arg[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] = threadIdx.x * 1e-3 + i; arg[i] = 0; }
    double best[8];
#pragma unroll
    for (int i = 0; i < 8; i++) best[i] = -1e300;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) {
                acc[i] = acc[i] + a;  // DADD
            } else {
                double v = acc[i] + a * it;  // stand-in for delta + logA: DFMA here, one fp64 op
                if (v > best[i]) { best[i] = v; arg[i] = it; }
                acc[i] = v * 0.5;  // keeps the chain alive: second... counted below
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i] + best[i] + arg[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the real Viterbi inner step: v = d[i] + a[i]; if (v > best) {best = v; arg = i;}
__global__ void __launch_bounds__(256) k_argmax(double *out, const double *din) {
    double d[16], a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { d[i] = din[i] + threadIdx.x; a[i] = din[16 + i]; }
    double tot = 0;
    int targ = 0;
    for (int it = 0; it < ITERS / 4; it++) {
        double best[4];
        int arg[4];
#pragma unroll
        for (int c = 0; c < 4; c++) { best[c] = d[c] + a[c]; arg[c] = c; }
#pragma unroll
        for (int i = 4; i < 16; i++) {
            int c = i & 3;
            double v = d[i] + a[i];
            if (v > best[c]) { best[c] = v; arg[c] = i; }
        }
#pragma unroll
        for (int c = 1; c < 4; c++)
            if (best[c] > best[0] || (best[c] == best[0] && arg[c] < arg[0])) { best[0] = best[c]; arg[0] = arg[c]; }
        tot += best[0];
        targ += arg[0];
#pragma unroll
        for (int i = 0; i < 16; i++) d[i] += 1e-3 * (arg[0] + 1);  // keep it live and data dependent
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot + targ;
}

__global__ void k_copy(const float4 *__restrict__ in, float4 *__restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

template <class F>
static float time_ms(F f, int reps = 5) {
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

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int grid = sms * 8, block = 256;
    float *fo;
    double *dout, *din;
    cudaMalloc(&fo, (size_t)grid * block * 4);
    cudaMalloc(&dout, (size_t)grid * block * 8);
    cudaMalloc(&din, 32 * 8);
    double h[32];
    for (int i = 0; i < 32; i++) h[i] = -1.0 * i * 0.37;
    cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
    double thr = (double)grid * block;
    float t0 = time_ms([&] { k_fp32<0><<<grid, block>>>(fo, 1.0001f, 0.5f); });
    float t1 = time_ms([&] { k_fp32<1><<<grid, block>>>(fo, 1.0001f, 0.5f); });
    float t2 = time_ms([&] { k_fp32<2><<<grid, block>>>(fo, 1.0001f, 0.5f); });
    float t3 = time_ms([&] { k_mufu<<<grid, block>>>(fo, 0.5f); });
    float t4 = time_ms([&] { k_fp64<0><<<grid, block>>>(dout, 1e-9); });
    float t5 = time_ms([&] { k_argmax<<<grid, block>>>(dout, din); });
    size_t n = (size_t)1 << 27;  // 2 GiB of float4
    float4 *ci, *co;
    cudaMalloc(&ci, n * 16);
    cudaMalloc(&co, n * 16);
    cudaMemset(ci, 1, n * 16);
    float t6 = time_ms([&] { k_copy<<<sms * 16, 512>>>(ci, co, n); });
    double fma = thr * ITERS * 16;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz_max\": %d,\n", p.name, sms, p.clockRate / 1000);
    printf(" \"fp32_ffma_scalar_tflops\": %.2f, \"fp32_ffma2_packed_tflops\": %.2f, \"fp32_mixed_tflops\": %.2f,\n",
           2 * fma / t0 * 1e-9, 2 * fma / t1 * 1e-9, 2 * fma / t2 * 1e-9);
    printf(" \"mufu_ex2_gops\": %.1f, \"fp64_dadd_gops\": %.1f,\n", thr * ITERS * 8 / t3 * 1e-6, thr * ITERS * 8 / t4 * 1e-6);
    printf(" \"fp64_viterbi_transitions_gps\": %.1f, \"hbm_copy_gbs\": %.1f}\n", thr * (ITERS / 4) * 16 / t5 * 1e-6,
           2.0 * n * 16 / t6 * 1e-6);
    return 0;
}
