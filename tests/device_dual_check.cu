// TEST INFRASTRUCTURE.  The expressions of tests/dual_check_expr.h evaluated
// in a KERNEL on the device (CUDA's libm, FMA contraction as in the product
// kernels): every elemental operation of include/gadfit_b200/dual.h through
// Passive, Dir2 (value, first and second directional derivative) and SDual
// (gradient).  tests/test_dual_device.py compares the result with the host
// build of the same header, with the reference's known answers
// (c++/tests/ad_forward_mode.cpp:16-256) and with the reference's own AdVar
// (tests/golden/dual_ops_ref.json).
#include "dual_check_expr.h"

#include <cuda_runtime.h>

namespace {

__global__ void k_dual_check(double* out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        dual_check::fill(out);
    }
}

// math_exp / math_rcp of dual.h beside CUDA's exp() and 1.0 / x
__global__ void k_math_check(const double* x, const int n, double* out)
{
    const int i { static_cast<int>(blockIdx.x * blockDim.x + threadIdx.x) };
    if (i < n) {
        out[i] = gadfit_b200::math_exp(x[i]);
        out[n + i] = ::exp(x[i]);
        out[2 * n + i] = gadfit_b200::math_rcp(x[i]);
        out[3 * n + i] = 1.0 / x[i];
    }
}

} // namespace

// out[0..n) = math_exp(x), [n..2n) = CUDA exp(x), [2n..3n) = math_rcp(x),
// [3n..4n) = 1.0 / x, all evaluated in a kernel
extern "C" __attribute__((visibility("default"))) int device_math_check(const double* x,
                                                                          int n,
                                                                          double* out)
{
    double* d_x { nullptr };
    double* d_out { nullptr };
    cudaError_t err { cudaMalloc(&d_x, n * sizeof(double)) };
    if (err == cudaSuccess) {
        err = cudaMalloc(&d_out, 4 * n * sizeof(double));
    }
    if (err == cudaSuccess) {
        err = cudaMemcpy(d_x, x, n * sizeof(double), cudaMemcpyHostToDevice);
    }
    if (err == cudaSuccess) {
        k_math_check<<<(n + 255) / 256, 256>>>(d_x, n, d_out);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) {
        err = cudaMemcpy(out, d_out, 4 * n * sizeof(double), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_x);
    cudaFree(d_out);
    return static_cast<int>(err);
}

// Returns 0 on success, a negative number if `cap` is too small, a CUDA error
// code otherwise; *n_entries / *width describe the layout of `out`.
extern "C" __attribute__((visibility("default"))) int device_dual_check(double* out,
                                                                          int cap,
                                                                          int* n_entries,
                                                                          int* width)
{
    constexpr int n { dual_check::n_entries * dual_check::width };
    *n_entries = dual_check::n_entries;
    *width = dual_check::width;
    if (cap < n) {
        return -1;
    }
    double* d_out { nullptr };
    cudaError_t err { cudaMalloc(&d_out, n * sizeof(double)) };
    if (err != cudaSuccess) {
        return static_cast<int>(err);
    }
    k_dual_check<<<1, 32>>>(d_out);
    err = cudaGetLastError();
    if (err == cudaSuccess) {
        err = cudaMemcpy(out, d_out, n * sizeof(double), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_out);
    return static_cast<int>(err);
}

extern "C" __attribute__((visibility("default"))) const char* device_dual_check_name(int i)
{
    return i >= 0 && i < dual_check::n_entries ? dual_check::names[i] : nullptr;
}
