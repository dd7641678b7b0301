// abi_common.cu -- error reporting, launch accounting and the FP issue-rate probe of libpdecg.so
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "pdecg_internal.h"

namespace pdecg {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PDECG_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return PDECG_OK;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Non-FMA floating-point issue-rate probe: 8 independent add/mul chains per thread.
template <typename T>
__global__ void __launch_bounds__(256) probe_fp_kernel(int iters, T* sink) {
    T x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = T(1) + T(threadIdx.x + c) * T(1e-3);
    const T m = T(0.999999), b = T(1e-6);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            x[c] = Strict<T>::mul(x[c], m);
            x[c] = Strict<T>::add(x[c], b);
        }
    }
    T s = T(0);
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c];
    if (s == T(-1)) sink[0] = s;   // never true; keeps the chains alive
}

template <typename T>
__global__ void constdiv_selftest_kernel(const T* __restrict__ x, T d, T* __restrict__ fast, T* __restrict__ ref, int64_t n) {
    const ConstDiv<T> cd(d);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        fast[i] = cd(x[i]);
        ref[i] = Strict<T>::div(x[i], d);
    }
}

template <typename T>
__global__ void fastdiv_selftest_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ fast, uint8_t* __restrict__ ok,
                                        T* __restrict__ ref, int64_t n) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        fast[i] = fast_div(a[i], b[i]);
        ok[i] = (DivWindow<T>::a(a[i]) & DivWindow<T>::b(b[i])) ? 1 : 0;
        ref[i] = Strict<T>::div(a[i], b[i]);
    }
}

__global__ void exp_selftest_kernel(const double* __restrict__ x, double* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) out[i] = exp_cr(x[i]);
}

}  // namespace pdecg

using namespace pdecg;

extern "C" int pdecg_selftest_exp(int64_t n, const double* x, double* out, pdecg_stream_t stream) {
    if (n < 0 || (n > 0 && (!x || !out))) return fail(PDECG_ERR_INVALID, "bad arguments");
    if (n == 0) return PDECG_OK;
    exp_selftest_kernel<<<592, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, n);
    count_launch();
    return check_launch("exp_selftest_kernel");
}

extern "C" int pdecg_selftest_constdiv(int32_t dtype, int64_t n, const void* x, double d, void* out_fast, void* out_ref,
                                       pdecg_stream_t stream) {
    if (n < 0 || (n > 0 && (!x || !out_fast || !out_ref))) return fail(PDECG_ERR_INVALID, "bad arguments");
    if (n == 0) return PDECG_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == PDECG_F64)
        constdiv_selftest_kernel<double><<<592, 256, 0, st>>>(static_cast<const double*>(x), d, static_cast<double*>(out_fast), static_cast<double*>(out_ref), n);
    else
        constdiv_selftest_kernel<float><<<592, 256, 0, st>>>(static_cast<const float*>(x), static_cast<float>(d), static_cast<float*>(out_fast), static_cast<float*>(out_ref), n);
    count_launch();
    return check_launch("constdiv_selftest_kernel");
}

extern "C" int pdecg_selftest_fastdiv(int32_t dtype, int64_t n, const void* a, const void* b, void* out_fast, uint8_t* out_ok,
                                      void* out_ref, pdecg_stream_t stream) {
    if (n < 0 || (n > 0 && (!a || !b || !out_fast || !out_ok || !out_ref))) return fail(PDECG_ERR_INVALID, "bad arguments");
    if (n == 0) return PDECG_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == PDECG_F64)
        fastdiv_selftest_kernel<double><<<592, 256, 0, st>>>(static_cast<const double*>(a), static_cast<const double*>(b), static_cast<double*>(out_fast), out_ok, static_cast<double*>(out_ref), n);
    else
        fastdiv_selftest_kernel<float><<<592, 256, 0, st>>>(static_cast<const float*>(a), static_cast<const float*>(b), static_cast<float*>(out_fast), out_ok, static_cast<float*>(out_ref), n);
    count_launch();
    return check_launch("fastdiv_selftest_kernel");
}

extern "C" const char* pdecg_last_error(void) { return g_err; }
extern "C" int32_t pdecg_abi_version(void) { return PDECG_ABI_VERSION; }
extern "C" int64_t pdecg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int pdecg_probe_fp_issue(int32_t dtype, int32_t blocks, int32_t threads, int32_t iters, float* ms,
                                    double* total_ops, pdecg_stream_t stream) {
    if (!ms || !total_ops) return fail(PDECG_ERR_INVALID, "null output");
    if (blocks <= 0 || threads <= 0 || threads > 256 || iters <= 0) return fail(PDECG_ERR_INVALID, "bad probe geometry");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void* sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, 16);
    if (e != cudaSuccess) return fail(PDECG_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {   // first pass warms up
        cudaEventRecord(e0, st);
        if (dtype == PDECG_F64) probe_fp_kernel<double><<<blocks, threads, 0, st>>>(iters, static_cast<double*>(sink));
        else probe_fp_kernel<float><<<blocks, threads, 0, st>>>(iters, static_cast<float*>(sink));
        cudaEventRecord(e1, st);
        count_launch();
    }
    e = cudaEventSynchronize(e1);
    int rc = PDECG_OK;
    if (e != cudaSuccess) rc = fail(PDECG_ERR_CUDA, "probe: %s", cudaGetErrorString(e));
    else {
        cudaEventElapsedTime(ms, e0, e1);
        *total_ops = 16.0 * iters * static_cast<double>(blocks) * threads;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return rc;
}
