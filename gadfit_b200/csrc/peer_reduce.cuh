// One-shot all-reduce of a few hundred doubles over NVLink / NVSwitch peer
// memory (sm_100a, one process per GPU of one box).
//
// The only data that crosses GPUs in the LM loop is the packed shared block
// [G, b_g], the Schur sums and chi2: at most p_g^2/2 + 2 p_g doubles (230 for
// 20 shared parameters).  That is pure latency; NCCL costs ~20-30 us per call
// for it.  Here every rank owns a MAILBOX in its HBM that all peers have
// mapped (cudaIpc): a rank stores its contribution directly into every
// peer's mailbox over NVLink, publishes a sequence number with a
// system-scope release store, waits (acquire) until all contributions have
// arrived in its own mailbox and sums them in RANK ORDER -- so every rank
// computes bit-identical sums, which is what lets the replicated shared solve
// and accept/reject decision stay in lockstep without a broadcast.
//
// Mailboxes are double buffered by the parity of the sequence number: a rank
// can be at most one collective ahead of the slowest rank (it needs that
// rank's contribution to finish its own), so two buffers suffice.
//
// The reference's counterpart is co_sum of the dense n x n matrix through
// OpenCoarrays/MPI (fortran/gadfit/gadfit.F90:700-701,735,1032).
#pragma once

#include "block_arrow_kernels.cuh"
#include "device_util.cuh"

namespace gadfit_b200 {

constexpr int kPeerMaxRanks { 16 };
constexpr int kPeerMaxDoubles { 1024 };

struct PeerMailbox
{
    unsigned long long flags[2][kPeerMaxRanks];
    double data[2][kPeerMaxRanks][kPeerMaxDoubles];
};

struct PeerTable
{
    PeerMailbox* box[kPeerMaxRanks]; // box[q]: mailbox in rank q's HBM
    int rank;
    int world;
    // A rank that never joins the collective (it failed before launching its
    // side) must not leave the others spinning for ever: after timeout_ns the
    // wait gives up and raises kFlagPeerTimeout in *status (LmState::flags),
    // which fit() reports as GADFIT_B200_NCCL_ERROR
    unsigned long long timeout_ns;
    unsigned* status;
};

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, const unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Block-wide: buf[0..n) -- global memory this block may just have written --
// is replaced by the sum over ranks.  Every thread of the block calls it.
__device__ __forceinline__ void
peer_all_reduce_block(const PeerTable& t, double* __restrict__ buf, const int n, const unsigned long long seq)
{
    const int parity { static_cast<int>(seq & 1ULL) };
    const int tid { static_cast<int>(threadIdx.x) };
    const int nthreads { static_cast<int>(blockDim.x) };
    __syncthreads();
    // 1. my contribution -> every rank's mailbox (peer stores over NVLink)
    for (int q = 0; q < t.world; ++q) {
        double* dst { t.box[q]->data[parity][t.rank] };
        for (int i = tid; i < n; i += nthreads) {
            dst[i] = buf[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish: flag[rank] = seq in every mailbox
    if (tid < t.world) {
        st_release_sys(&t.box[tid]->flags[parity][t.rank], seq);
    }
    // 3. wait until every rank's contribution has landed in MY mailbox
    if (tid < t.world) {
        const unsigned long long* flag { &t.box[t.rank]->flags[parity][tid] };
        const unsigned long long begin { global_timer_ns() };
        while (ld_acquire_sys(flag) < seq) {
            if (global_timer_ns() - begin > t.timeout_ns) {
                if (t.status != nullptr) {
                    atomicOr(t.status, kFlagPeerTimeout);
                }
                break;
            }
        }
    }
    __syncthreads();
    // 4. sum in rank order (identical on every rank); bypass L1, the lines
    //    were written by peers
    const PeerMailbox* mine { t.box[t.rank] };
    for (int i = tid; i < n; i += nthreads) {
        double sum { 0.0 };
        for (int q = 0; q < t.world; ++q) {
            sum += __ldcv(&mine->data[parity][q][i]);
        }
        buf[i] = sum;
    }
}

// One CTA.  buf[0..n) is replaced by the sum over ranks.
static __global__ void __launch_bounds__(256)
  k_peer_all_reduce(const PeerTable t, double* __restrict__ buf, const int n, const unsigned long long seq)
{
    peer_all_reduce_block(t, buf, n, seq);
}

// Reduction + collective in ONE kernel: the last stage of a deterministic
// local reduction hands its result straight to the peers' mailboxes, so no
// launch and no HBM round trip sits between the math and the exchange.
//   chi2 sweep: sum of the CTA partials (k_sum_partials) + all-reduce
static __global__ void __launch_bounds__(kThreads)
  k_sum_partials_peer(const double* __restrict__ partials,
                      const int n,
                      double* __restrict__ out,
                      const PeerTable t,
                      const unsigned long long seq)
{
    __shared__ double s_red[kThreads / 32];
    double sum { 0.0 };
    for (int i = static_cast<int>(threadIdx.x); i < n; i += kThreads) {
        sum += partials[i];
    }
    const double total { block_sum(sum, s_red) };
    if (threadIdx.x == 0) {
        *out = total;
    }
    peer_all_reduce_block(t, out, 1, seq);
}
//   shared block / Schur sums: last stage of k_column_sums + all-reduce
static __global__ void __launch_bounds__(256)
  k_column_sums_peer(const double* __restrict__ in,
                     const int n_rows,
                     const int V,
                     double* __restrict__ out,
                     const PeerTable t,
                     const unsigned long long seq)
{
    for (int v = static_cast<int>(threadIdx.x); v < V; v += blockDim.x) {
        double sum { 0.0 };
        for (int r = 0; r < n_rows; ++r) {
            sum += in[static_cast<long long>(r) * V + v];
        }
        out[v] = sum;
    }
    peer_all_reduce_block(t, out, V, seq);
}

// Host side: mailbox allocation and the exchange of IPC handles (through the
// NCCL communicator that exists anyway for the large, rare reductions)
struct PeerReduce
{
    PeerTable table {};
    PeerMailbox* local { nullptr };
    void* opened[kPeerMaxRanks] {};
    unsigned long long seq { 0 };
    bool active { false };

    ~PeerReduce() { release(); }

    auto release() -> void
    {
        for (auto& p : opened) {
            if (p != nullptr) {
                cudaIpcCloseMemHandle(p);
                p = nullptr;
            }
        }
        if (local != nullptr) {
            cudaFree(local);
            local = nullptr;
        }
        active = false;
        seq = 0;
    }

    // Collective over `comm`.  Returns false (and stays inactive) if peer
    // mapping is not possible; the caller then keeps using NCCL.
    auto setup(const int rank, const int world, ncclComm_t comm, cudaStream_t stream) -> bool
    {
        release();
        if (world < 2 || world > kPeerMaxRanks || std::getenv("GADFIT_B200_NO_PEER_REDUCE") != nullptr) {
            return false;
        }
        auto& nccl { Nccl::instance() };
        if (nccl.allGather == nullptr) {
            return false;
        }
        GADFIT_CUDA_TRY(cudaMalloc(&local, sizeof(PeerMailbox)));
        GADFIT_CUDA_TRY(cudaMemsetAsync(local, 0, sizeof(PeerMailbox), stream));
        cudaIpcMemHandle_t mine {};
        int ok { cudaIpcGetMemHandle(&mine, local) == cudaSuccess ? 1 : 0 };
        (void)cudaGetLastError();
        // gather [ok flag (8 bytes) | handle (64 bytes)] from every rank
        constexpr size_t record { 8 + sizeof(cudaIpcMemHandle_t) };
        std::vector<unsigned char> send(record, 0);
        std::vector<unsigned char> recv(record * world, 0);
        send[0] = static_cast<unsigned char>(ok);
        std::memcpy(send.data() + 8, &mine, sizeof(mine));
        DeviceBuffer<unsigned char> d_send;
        DeviceBuffer<unsigned char> d_recv;
        d_send.upload(send, stream);
        d_recv.ensure(record * world);
        nccl.check(nccl.allGather(d_send.ptr, d_recv.ptr, record, ncclUint8, comm, stream),
                   "ncclAllGather");
        GADFIT_CUDA_TRY(cudaMemcpyAsync(
          recv.data(), d_recv.ptr, record * world, cudaMemcpyDeviceToHost, stream));
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
        bool all_ok { true };
        for (int q = 0; q < world; ++q) {
            all_ok = all_ok && recv[record * q] == 1;
        }
        table.rank = rank;
        table.world = world;
        // GADFIT_B200_PEER_TIMEOUT_S: seconds a rank waits for its peers (default 30)
        const char* timeout_env { std::getenv("GADFIT_B200_PEER_TIMEOUT_S") };
        const double timeout_s { timeout_env != nullptr ? std::atof(timeout_env) : 30.0 };
        table.timeout_ns = static_cast<unsigned long long>((timeout_s > 0.0 ? timeout_s : 30.0) * 1e9);
        for (int q = 0; q < world && all_ok; ++q) {
            if (q == rank) {
                table.box[q] = local;
                continue;
            }
            cudaIpcMemHandle_t theirs {};
            std::memcpy(&theirs, recv.data() + record * q + 8, sizeof(theirs));
            void* p { nullptr };
            if (cudaIpcOpenMemHandle(&p, theirs, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                (void)cudaGetLastError();
                all_ok = false;
                break;
            }
            opened[q] = p;
            table.box[q] = static_cast<PeerMailbox*>(p);
        }
        // every rank must agree on the outcome: all-reduce the verdict
        DeviceBuffer<double> verdict;
        const double mine_ok { all_ok ? 0.0 : 1.0 };
        verdict.upload(&mine_ok, 1, stream);
        nccl.check(nccl.allReduce(verdict.ptr, verdict.ptr, 1, ncclDouble, ncclSum, comm, stream),
                   "ncclAllReduce");
        double failures { 0.0 };
        GADFIT_CUDA_TRY(cudaMemcpyAsync(
          &failures, verdict.ptr, sizeof(double), cudaMemcpyDeviceToHost, stream));
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
        if (failures != 0.0) {
            release();
            return false;
        }
        active = true;
        return true;
    }

    auto all_reduce(double* buf, const int n, cudaStream_t stream) -> void
    {
        ++seq;
        k_peer_all_reduce<<<1, 256, 0, stream>>>(table, buf, n, seq);
    }
    // fused with the last stage of the local reduction
    auto sum_partials_all_reduce(const double* partials, const int n, double* out,
                                 cudaStream_t stream) -> void
    {
        ++seq;
        k_sum_partials_peer<<<1, kThreads, 0, stream>>>(partials, n, out, table, seq);
    }
    auto column_sums_all_reduce(const double* in, const int n_rows, const int V, double* out,
                                const int threads, cudaStream_t stream) -> void
    {
        ++seq;
        k_column_sums_peer<<<1, threads, 0, stream>>>(in, n_rows, V, out, table, seq);
    }
};

} // namespace gadfit_b200
