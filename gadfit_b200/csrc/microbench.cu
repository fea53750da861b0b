// FP64 throughput micro-benchmarks: the roofline denominator of bench.py and
// the evidence behind the choice between the DFMA vector pipe and the DMMA
// tensor path for the normal-equation accumulation.
#include "device_util.cuh"

namespace gadfit_b200 {
namespace {

constexpr int kChains { 8 };

// Register-resident DFMA chains: 2 * kChains * iters flops per thread
__global__ void k_dfma(double* out, const int iters, const double a, const double b)
{
    double acc[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        acc[c] = threadIdx.x + c;
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            acc[c] = ::fma(acc[c], a, b);
        }
    }
    double sum { 0.0 };
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        sum += acc[c];
    }
    if (sum == 12345.678) {
        out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    }
}

__device__ __forceinline__ void dmma884(double (&c)[2], const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// DMMA m8n8k4 chains: 512 flops per warp-level mma
__global__ void k_dmma(double* out, const int iters, const double a, const double b)
{
    double acc[kChains][2];
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        acc[c][0] = threadIdx.x + c;
        acc[c][1] = c;
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            dmma884(acc[c], a, b);
        }
    }
    double sum { 0.0 };
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        sum += acc[c][0] + acc[c][1];
    }
    if (sum == 12345.678) {
        out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    }
}

// N independent DMMA accumulator chains in one warp (one warp per SM
// sub-partition): how much of the pipe can a single warp keep busy?
template <int N>
__global__ void k_dmma_chains(double* out, const int iters, const double a, const double b)
{
    double acc[N][2];
#pragma unroll
    for (int c = 0; c < N; ++c) {
        acc[c][0] = threadIdx.x + c;
        acc[c][1] = c;
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < N; ++c) {
            dmma884(acc[c], a, b);
        }
    }
    double sum { 0.0 };
#pragma unroll
    for (int c = 0; c < N; ++c) {
        sum += acc[c][0] + acc[c][1];
    }
    if (sum == 12345.678) {
        out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    }
}

__global__ void k_dfma_chain(double* out, const int iters, const double a, const double b)
{
    double acc { static_cast<double>(threadIdx.x) };
    for (int i = 0; i < iters; ++i) {
        acc = ::fma(acc, a, b);
    }
    if (acc == 12345.678) {
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    }
}

// Both in the same warp, interleaved: do the two share one pipe?
__global__ void k_mixed(double* out, const int iters, const double a, const double b)
{
    double acc[kChains / 2][2];
    double f[kChains / 2 * 8];
#pragma unroll
    for (int c = 0; c < kChains / 2; ++c) {
        acc[c][0] = threadIdx.x + c;
        acc[c][1] = c;
    }
#pragma unroll
    for (int c = 0; c < kChains / 2 * 8; ++c) {
        f[c] = threadIdx.x - c;
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains / 2; ++c) {
            dmma884(acc[c], a, b);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                f[c * 8 + k] = ::fma(f[c * 8 + k], a, b);
            }
        }
    }
    double sum { 0.0 };
#pragma unroll
    for (int c = 0; c < kChains / 2; ++c) {
        sum += acc[c][0] + acc[c][1];
    }
#pragma unroll
    for (int c = 0; c < kChains / 2 * 8; ++c) {
        sum += f[c];
    }
    if (sum == 12345.678) {
        out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    }
}

} // namespace

#pragma GCC visibility push(default)
extern "C" {

// which: 0 = DFMA, 1 = DMMA m8n8k4, 2 = DMMA + DFMA interleaved (equal flops
// of each); *out = FLOP/s (best of 5)
int gadfit_b200_microbench(int which, double* out)
{
    try {
        int device { 0 };
        int sms { 0 };
        GADFIT_CUDA_TRY(cudaGetDevice(&device));
        GADFIT_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        constexpr int threads { 256 };
        const int blocks { sms * 8 };
        constexpr int iters { 1 << 14 };
        DeviceBuffer<double> sink;
        sink.ensure(static_cast<size_t>(blocks) * threads);
        cudaEvent_t e0 {};
        cudaEvent_t e1 {};
        GADFIT_CUDA_TRY(cudaEventCreate(&e0));
        GADFIT_CUDA_TRY(cudaEventCreate(&e1));
        double best { 0.0 };
        for (int rep = 0; rep < 6; ++rep) {
            GADFIT_CUDA_TRY(cudaEventRecord(e0));
            double flops { 0.0 };
            const double n_threads { static_cast<double>(blocks) * threads };
            switch (which) {
            case 0:
                k_dfma<<<blocks, threads>>>(sink.ptr, iters, 1.0000001, 1e-9);
                flops = n_threads * iters * kChains * 2.0;
                break;
            case 1:
                k_dmma<<<blocks, threads>>>(sink.ptr, iters, 1.0000001, 1e-9);
                flops = n_threads / 32.0 * iters * kChains * 512.0;
                break;
            case 2:
                k_mixed<<<blocks, threads>>>(sink.ptr, iters, 1.0000001, 1e-9);
                flops = n_threads / 32.0 * iters * (kChains / 2) * 512.0
                        + n_threads * iters * (kChains / 2) * 8 * 2.0;
                break;
            case 3: // DMMA throughput of ONE warp per SM sub-partition with 6 chains
            case 4: // ... with 1 chain (= issue-to-issue latency of a dependent DMMA)
            case 5: // DFMA, one warp per sub-partition, 1 chain (latency)
                if (which == 3) {
                    k_dmma_chains<6><<<sms, 128>>>(sink.ptr, iters, 1.0000001, 1e-9);
                    flops = sms * 4.0 * iters * 6 * 512.0;
                } else if (which == 4) {
                    k_dmma_chains<1><<<sms, 128>>>(sink.ptr, iters, 1.0000001, 1e-9);
                    flops = sms * 4.0 * iters * 1 * 512.0;
                } else {
                    k_dfma_chain<<<sms, 128>>>(sink.ptr, iters, 1.0000001, 1e-9);
                    flops = sms * 128.0 * iters * 2.0;
                }
                break;
            default:
                return GADFIT_B200_INVALID_ARGUMENT;
            }
            GADFIT_CUDA_TRY(cudaGetLastError());
            GADFIT_CUDA_TRY(cudaEventRecord(e1));
            GADFIT_CUDA_TRY(cudaEventSynchronize(e1));
            float ms { 0.0F };
            GADFIT_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0) {
                best = std::max(best, flops / (static_cast<double>(ms) * 1e-3));
            }
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *out = best;
        return GADFIT_B200_OK;
    } catch (const std::exception&) {
        return GADFIT_B200_CUDA_ERROR;
    }
}

int gadfit_b200_measure_fp64_peak(double* flops_per_second)
{
    return gadfit_b200_microbench(0, flops_per_second);
}

} // extern "C"
#pragma GCC visibility pop

} // namespace gadfit_b200
