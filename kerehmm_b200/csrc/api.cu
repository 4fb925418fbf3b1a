// Error plumbing, device info and small utility kernels of libkhmm.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace khmm {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return KHMM_ERR_CUDA;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        cached = p.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

static const char *kSwitchNames[SW_COUNT] = {"KHMM_RC_MODE", "KHMM_RC_BN", "KHMM_VIT_SCREEN", "KHMM_BW_BACKWARD",
                                             "KHMM_SMALL_MMA"};
static int g_switch[SW_COUNT];
static bool g_switch_init = false;

static void switches_init() {
    if (g_switch_init) return;
    for (int k = 0; k < SW_COUNT; k++) {
        const char *e = getenv(kSwitchNames[k]);
        g_switch[k] = (e && *e) ? atoi(e) : -1;
    }
    g_switch_init = true;
}

int dev_switch(DevSwitch which) {
    switches_init();
    return g_switch[which];
}

template <typename R>
__global__ void transpose_kernel(int rows, int cols, const R *__restrict__ in, R *__restrict__ out) {
    __shared__ R tile[32][33];
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        int r = r0 + dy, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[dy][threadIdx.x] = in[(int64_t)r * cols + c];
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        int c = c0 + dy, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(int64_t)c * rows + r] = tile[threadIdx.x][dy];
    }
}

template <typename R>
static int transpose_impl(int rows, int cols, const R *in, R *out, khmm_stream_t stream) {
    KHMM_REQUIRE(rows > 0 && cols > 0 && in && out, "khmm_transpose: bad arguments");
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_kernel<R><<<grid, block, 0, (cudaStream_t)stream>>>(rows, cols, in, out);
    KHMM_LAUNCH_CHECK("transpose_kernel");
    return KHMM_OK;
}

}  // namespace khmm

extern "C" {

const char *khmm_last_error(void) { return khmm::g_err; }

int khmm_version(void) { return 200; }

int khmm_debug_switch(const char *name, int value) {
    KHMM_REQUIRE(name, "debug_switch: NULL name");
    khmm::switches_init();
    for (int k = 0; k < khmm::SW_COUNT; k++)
        if (strcmp(name, khmm::kSwitchNames[k]) == 0) {
            khmm::g_switch[k] = value;
            return KHMM_OK;
        }
    khmm::set_error("debug_switch: unknown switch %s", name);
    return KHMM_ERR_INVALID;
}

int khmm_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem) {
    cudaDeviceProp p;
    KHMM_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return KHMM_OK;
}

int khmm_transpose_f32(int rows, int cols, const float *in, float *out, khmm_stream_t stream) {
    return khmm::transpose_impl<float>(rows, cols, in, out, stream);
}
int khmm_transpose_f64(int rows, int cols, const double *in, double *out, khmm_stream_t stream) {
    return khmm::transpose_impl<double>(rows, cols, in, out, stream);
}

}  // extern "C"
