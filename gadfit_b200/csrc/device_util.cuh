// Small host-side helpers around the CUDA runtime and NCCL (loaded lazily).
#pragma once

#include "gadfit_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h> // types and enums only; the functions are resolved with dlsym

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace gadfit_b200 {

// Carries a gadfit_b200_status across the C++ internals to the C-ABI boundary
struct StatusError : std::runtime_error
{
    int status;
    StatusError(const int status, const std::string& msg)
      : std::runtime_error { msg }, status { status }
    {}
};

#define GADFIT_CUDA_TRY(expr)                                                  \
    do {                                                                       \
        const cudaError_t err_ { (expr) };                                     \
        if (err_ != cudaSuccess) {                                             \
            throw ::gadfit_b200::StatusError {                                 \
                GADFIT_B200_CUDA_ERROR,                                        \
                std::string { #expr } + ": " + cudaGetErrorString(err_)        \
            };                                                                 \
        }                                                                      \
    } while (0)

template <class T>
struct DeviceBuffer
{
    T* ptr { nullptr };
    size_t capacity { 0 };

    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer&) = delete;
    auto operator=(const DeviceBuffer&) -> DeviceBuffer& = delete;
    ~DeviceBuffer() { release(); }

    auto release() -> void
    {
        if (ptr != nullptr) {
            cudaFree(ptr);
            ptr = nullptr;
            capacity = 0;
        }
    }
    // Grow-only allocation; contents are NOT preserved
    auto ensure(const size_t n) -> T*
    {
        if (n > capacity) {
            release();
            GADFIT_CUDA_TRY(cudaMalloc(&ptr, (n > 0 ? n : 1) * sizeof(T)));
            capacity = n;
        }
        return ptr;
    }
    auto upload(const T* host, const size_t n, cudaStream_t stream) -> void
    {
        ensure(n);
        if (n > 0) {
            GADFIT_CUDA_TRY(cudaMemcpyAsync(
              ptr, host, n * sizeof(T), cudaMemcpyHostToDevice, stream));
        }
    }
    auto upload(const std::vector<T>& host, cudaStream_t stream) -> void
    {
        upload(host.data(), host.size(), stream);
    }
    auto zero(const size_t n, cudaStream_t stream) -> void
    {
        ensure(n);
        if (n > 0) {
            GADFIT_CUDA_TRY(cudaMemsetAsync(ptr, 0, n * sizeof(T), stream));
        }
    }
};

// Device-time phase timer: events are recorded on the solver's stream and
// only read after the fit has finished, so timing never adds a sync.
struct PhaseTimer
{
    std::vector<cudaEvent_t> events; // start/stop pairs
    size_t used { 0 };
    double seconds { 0.0 };
    int calls { 0 };

    ~PhaseTimer()
    {
        for (auto e : events) {
            cudaEventDestroy(e);
        }
    }
    auto reset() -> void
    {
        used = 0;
        seconds = 0.0;
        calls = 0;
    }
    auto next() -> cudaEvent_t
    {
        if (used == events.size()) {
            cudaEvent_t e {};
            GADFIT_CUDA_TRY(cudaEventCreate(&e));
            events.push_back(e);
        }
        return events[used++];
    }
    // false: count the calls only (no events on the stream)
    bool enabled { true };
    auto start(cudaStream_t stream) -> void
    {
        if (enabled) {
            GADFIT_CUDA_TRY(cudaEventRecord(next(), stream));
        }
    }
    auto stop(cudaStream_t stream) -> void
    {
        if (enabled) {
            GADFIT_CUDA_TRY(cudaEventRecord(next(), stream));
        }
        ++calls;
    }
    // Call after the stream has been synchronised
    auto collect() -> void
    {
        seconds = 0.0;
        for (size_t i = 0; i + 1 < used; i += 2) {
            float ms { 0.0F };
            GADFIT_CUDA_TRY(cudaEventElapsedTime(&ms, events[i], events[i + 1]));
            seconds += static_cast<double>(ms) * 1e-3;
        }
    }
};

// NCCL entry points resolved at run time: the library has no link-time
// dependency on NCCL, and inside a PyTorch process it binds to the libnccl
// torch has already loaded.
struct Nccl
{
    using GetUniqueId = ncclResult_t (*)(ncclUniqueId*);
    using CommInitRank = ncclResult_t (*)(ncclComm_t*, int, ncclUniqueId, int);
    using CommDestroy = ncclResult_t (*)(ncclComm_t);
    using AllReduce = ncclResult_t (*)(const void*,
                                       void*,
                                       size_t,
                                       ncclDataType_t,
                                       ncclRedOp_t,
                                       ncclComm_t,
                                       cudaStream_t);
    using GetErrorString = const char* (*)(ncclResult_t);

    void* handle { nullptr };
    GetUniqueId getUniqueId { nullptr };
    CommInitRank commInitRank { nullptr };
    CommDestroy commDestroy { nullptr };
    AllReduce allReduce { nullptr };
    using AllGather = ncclResult_t (*)(const void*,
                                       void*,
                                       size_t,
                                       ncclDataType_t,
                                       ncclComm_t,
                                       cudaStream_t);
    AllGather allGather { nullptr }; // optional: only the peer-mailbox set-up needs it
    GetErrorString getErrorString { nullptr };

    static auto instance() -> Nccl&
    {
        static Nccl nccl {};
        if (nccl.handle == nullptr) {
            nccl.load();
        }
        return nccl;
    }

    auto load() -> void
    {
        const char* override_path { std::getenv("GADFIT_B200_NCCL") };
        const char* candidates[] { override_path, "libnccl.so.2", "libnccl.so" };
        for (const char* name : candidates) {
            if (name == nullptr) {
                continue;
            }
            handle = dlopen(name, RTLD_NOW | RTLD_NOLOAD);
            if (handle == nullptr) {
                handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            }
            if (handle != nullptr) {
                break;
            }
        }
        if (handle == nullptr) {
            throw StatusError { GADFIT_B200_NCCL_ERROR,
                                std::string { "cannot load libnccl: " } + dlerror() };
        }
        getUniqueId = reinterpret_cast<GetUniqueId>(dlsym(handle, "ncclGetUniqueId"));
        commInitRank =
          reinterpret_cast<CommInitRank>(dlsym(handle, "ncclCommInitRank"));
        commDestroy = reinterpret_cast<CommDestroy>(dlsym(handle, "ncclCommDestroy"));
        allReduce = reinterpret_cast<AllReduce>(dlsym(handle, "ncclAllReduce"));
        allGather = reinterpret_cast<AllGather>(dlsym(handle, "ncclAllGather"));
        getErrorString =
          reinterpret_cast<GetErrorString>(dlsym(handle, "ncclGetErrorString"));
        if (getUniqueId == nullptr || commInitRank == nullptr
            || commDestroy == nullptr || allReduce == nullptr
            || getErrorString == nullptr) {
            throw StatusError { GADFIT_B200_NCCL_ERROR,
                                "libnccl lacks a required symbol" };
        }
    }

    auto check(const ncclResult_t r, const char* what) const -> void
    {
        if (r != ncclSuccess) {
            throw StatusError { GADFIT_B200_NCCL_ERROR,
                                std::string { what } + ": " + getErrorString(r) };
        }
    }
};

} // namespace gadfit_b200
