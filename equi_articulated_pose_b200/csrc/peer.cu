// peer.cu -- SyncBatchNorm exchange over NVLink peer memory (one process per GPU, one node).
//
// The reference converts every BatchNorm to nn.SyncBatchNorm under DDP
// (SPConvNets/trainer_unsup_arti_align.py:430): 14 layers x (forward + backward) tiny all-reduces per
// step (2*C fp64 sums + a row count).  An NCCL all-reduce of 4 KB costs ~25-30 us of launch + protocol latency;
// here the exchange is ONE single-CTA kernel per reduction: every rank stores its sums straight into a mailbox
// slot in each peer's HBM (cudaIpc-mapped, NVLink 5 / NVSwitch stores), publishes a sequence flag with a
// system-scope release store, spins on the flags of its own mailbox and adds the slots in rank order (so every
// rank gets bit-identical sums).  The forward variant also finalises (mean, invstd, running statistics) in the
// same kernel.
//
// Mailbox (per rank, written by its peers):  slot[s % 4][src rank][SLOT_DOUBLES]
//   doubles 0..n-1 = payload, word SLOT_DOUBLES-1 = sequence flag (uint64).  A rank can run at most one exchange
// ahead of its slowest peer (it needs that peer's contribution to finish), so 4 slots are never overwritten live.
// Behind the slots every rank keeps two private words: the exchange COUNTER (the sequence number lives on the device:
// the kernel takes the next value itself when the host passes seq = 0, so a captured CUDA graph can be replayed) and a
// STATUS word (set to the offending rank + 1 when a peer did not show up within the timeout; the kernel then returns
// with unusable results instead of trapping, and vgtkb_peer_status reports it).  The timeout defaults to 600 s
// (VGTKB_PEER_TIMEOUT_S), the order of a process-group timeout: a rank may lag behind for a long time legitimately
// (rank-0-only evaluation or checkpointing, a stalled data loader, first-step lazy initialisation).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace vgtkb {

constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_MAX_N = 2 * 1024 + 1;            // 2*C sums + row count, C <= 1024
constexpr int PEER_SLOT_DOUBLES = PEER_MAX_N + 1;   // + flag
constexpr int PEER_SLOTS = 4;
constexpr int PEER_THREADS = 256;

struct PeerBoxes {
    double* box[PEER_MAX_WORLD];
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ double* peer_slot(double* box, int world, unsigned long long seq, int src) {
    return box + ((size_t)(seq % PEER_SLOTS) * world + src) * PEER_SLOT_DOUBLES;
}
__host__ __device__ __forceinline__ size_t peer_tail_offset(int world) {   // doubles before the private words
    return (size_t)PEER_SLOTS * world * PEER_SLOT_DOUBLES;
}

// the sequence number of this exchange: the host's value, or (seq == 0) the next value of the device-side counter
__device__ unsigned long long peer_next_seq(unsigned long long seq, int rank, int world, const PeerBoxes& pb) {
    __shared__ unsigned long long s_seq;
    if (threadIdx.x == 0) {
        unsigned long long* ctr = reinterpret_cast<unsigned long long*>(pb.box[rank] + peer_tail_offset(world));
        if (seq == 0) {
            seq = *ctr + 1;
            *ctr = seq;
        } else {
            *ctr = seq;
        }
        s_seq = seq;
    }
    __syncthreads();
    return s_seq;
}

// push `local[0..n)` to every rank's mailbox, wait for everybody's contribution in mine
__device__ void peer_exchange(int n, const double* __restrict__ local, int rank, int world, const PeerBoxes& pb,
                              unsigned long long seq, unsigned long long timeout_ns) {
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const double v = local[t];
        for (int p = 0; p < world; ++p) st_relaxed_sys_f64(peer_slot(pb.box[p], world, seq, rank) + t, v);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world)
        st_release_sys_u64(reinterpret_cast<unsigned long long*>(peer_slot(pb.box[threadIdx.x], world, seq, rank) + PEER_SLOT_DOUBLES - 1), seq);
    if ((int)threadIdx.x < world) {
        const unsigned long long* flag =
            reinterpret_cast<const unsigned long long*>(peer_slot(pb.box[rank], world, seq, threadIdx.x) + PEER_SLOT_DOUBLES - 1);
        const unsigned long long t0 = global_timer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys_u64(flag) != seq) {
            if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > timeout_ns) {
                printf("vgtkb peer exchange: rank %d gave up waiting for rank %d (seq %llu); results are invalid\n", rank,
                       (int)threadIdx.x, seq);
                reinterpret_cast<unsigned long long*>(pb.box[rank] + peer_tail_offset(world))[1] = threadIdx.x + 1;   // status
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ double peer_sum(int t, int rank, int world, const PeerBoxes& pb, unsigned long long seq) {
    double s = 0.0;
    for (int p = 0; p < world; ++p) s += ld_relaxed_sys_f64(peer_slot(pb.box[rank], world, seq, p) + t);   // rank order
    return s;
}

__global__ void __launch_bounds__(PEER_THREADS)
peer_allreduce_kernel(int n, double* __restrict__ buf, int rank, int world, PeerBoxes pb, unsigned long long seq,
                      unsigned long long timeout_ns) {
    seq = peer_next_seq(seq, rank, world, pb);
    peer_exchange(n, buf, rank, world, pb, seq, timeout_ns);
    for (int t = threadIdx.x; t < n; t += blockDim.x) buf[t] = peer_sum(t, rank, world, pb, seq);
}

// exchange of [sum x | sum x^2 | rows] + the statistics of vgtkb_norm_finalize, one kernel
__global__ void __launch_bounds__(PEER_THREADS)
norm_finalize_peer_kernel(int c, float eps, double* __restrict__ scratch, float* __restrict__ stats, float* running_mean,
                          float* running_var, float momentum, int rank, int world, PeerBoxes pb, unsigned long long seq,
                          unsigned long long timeout_ns) {
    seq = peer_next_seq(seq, rank, world, pb);
    peer_exchange(2 * c + 1, scratch, rank, world, pb, seq, timeout_ns);
    const double n = peer_sum(2 * c, rank, world, pb, seq);
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        const double s1 = peer_sum(ch, rank, world, pb, seq), s2 = peer_sum(c + ch, rank, world, pb, seq);
        const double mean = s1 / n;
        double var = s2 / n - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[ch] = (float)mean;
        stats[c + ch] = (float)(1.0 / sqrt(var + (double)eps));
        if (running_mean != nullptr) {
            const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
            running_mean[ch] = (float)((1.0 - momentum) * running_mean[ch] + momentum * mean);
            running_var[ch] = (float)((1.0 - momentum) * running_var[ch] + momentum * unbiased);
        }
        scratch[ch] = s1;          // global sums, as after an all-reduce
        scratch[c + ch] = s2;
    }
    if (threadIdx.x == 0) scratch[2 * c] = n;
}

static unsigned long long peer_timeout_ns() {
    static const unsigned long long v = []() {
        const char* e = getenv("VGTKB_PEER_TIMEOUT_S");
        const double sec = e != nullptr && atof(e) > 0.0 ? atof(e) : 600.0;
        return (unsigned long long)(sec * 1e9);
    }();
    return v;
}

static int fill_boxes(PeerBoxes& pb, int rank, int world, void* const* mailboxes) {
    VGTKB_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "peer: bad rank/world");
    VGTKB_REQUIRE(mailboxes != nullptr, "peer: no mailboxes");
    for (int p = 0; p < world; ++p) {
        VGTKB_REQUIRE(mailboxes[p] != nullptr, "peer: mailbox of a rank is NULL");
        pb.box[p] = static_cast<double*>(mailboxes[p]);
    }
    return VGTKB_OK;
}

}  // namespace vgtkb

using namespace vgtkb;

extern "C" int vgtkb_peer_mailbox_bytes(int world, int64_t* bytes) {
    VGTKB_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && bytes != nullptr, "peer_mailbox_bytes: bad argument");
    *bytes = ((int64_t)PEER_SLOTS * world * PEER_SLOT_DOUBLES + 2) * (int64_t)sizeof(double);   // + counter, status
    return VGTKB_OK;
}

extern "C" int vgtkb_peer_alloc(int64_t bytes, void** dev_ptr, void* ipc_handle) {
    VGTKB_REQUIRE(bytes > 0 && dev_ptr != nullptr && ipc_handle != nullptr, "peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == VGTKB_IPC_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    VGTKB_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(ipc_handle), p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("peer_alloc: %s", cudaGetErrorString(e));
        return VGTKB_ECUDA;
    }
    *dev_ptr = p;
    return VGTKB_OK;
}

extern "C" int vgtkb_peer_open(const void* ipc_handle, void** dev_ptr) {
    VGTKB_REQUIRE(ipc_handle != nullptr && dev_ptr != nullptr, "peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    VGTKB_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return VGTKB_OK;
}

extern "C" int vgtkb_peer_close(void* dev_ptr) {
    if (dev_ptr != nullptr) VGTKB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return VGTKB_OK;
}

extern "C" int vgtkb_peer_free(void* dev_ptr) {
    if (dev_ptr != nullptr) VGTKB_CUDA(cudaFree(dev_ptr));
    return VGTKB_OK;
}

extern "C" int vgtkb_peer_allreduce_f64(int n, double* buf, int rank, int world, void* const* mailboxes, uint64_t seq,
                                        void* stream) {
    VGTKB_REQUIRE(n > 0 && n <= PEER_MAX_N && buf != nullptr, "peer_allreduce_f64: bad argument");
    PeerBoxes pb;
    const int rc = fill_boxes(pb, rank, world, mailboxes);
    if (rc != VGTKB_OK) return rc;
    peer_allreduce_kernel<<<1, PEER_THREADS, 0, (cudaStream_t)stream>>>(n, buf, rank, world, pb, (unsigned long long)seq,
                                                                        peer_timeout_ns());
    return check_launch("peer_allreduce_f64");
}

extern "C" int vgtkb_norm_finalize_peer(int c, float eps, double* scratch, float* stats, float* running_mean,
                                        float* running_var, float momentum, int rank, int world, void* const* mailboxes,
                                        uint64_t seq, void* stream) {
    VGTKB_REQUIRE(c > 0 && 2 * c + 1 <= PEER_MAX_N && scratch != nullptr && stats != nullptr, "norm_finalize_peer: bad argument");
    PeerBoxes pb;
    const int rc = fill_boxes(pb, rank, world, mailboxes);
    if (rc != VGTKB_OK) return rc;
    norm_finalize_peer_kernel<<<1, PEER_THREADS, 0, (cudaStream_t)stream>>>(c, eps, scratch, stats, running_mean, running_var,
                                                                           momentum, rank, world, pb, (unsigned long long)seq,
                                                                           peer_timeout_ns());
    return check_launch("norm_finalize_peer");
}

// status word of this rank's mailbox (device -> host; synchronises the stream): 0 = every exchange completed,
// r + 1 = rank r did not arrive within the timeout in some exchange (the sums of that exchange are invalid)
extern "C" int vgtkb_peer_status(const void* own_mailbox, int world, int64_t* status, void* stream) {
    VGTKB_REQUIRE(own_mailbox != nullptr && status != nullptr && world >= 1 && world <= PEER_MAX_WORLD, "peer_status: bad argument");
    unsigned long long v = 0;
    const double* tail = static_cast<const double*>(own_mailbox) + peer_tail_offset(world);
    VGTKB_CUDA(cudaMemcpyAsync(&v, reinterpret_cast<const unsigned long long*>(tail) + 1, sizeof(v), cudaMemcpyDeviceToHost,
                               (cudaStream_t)stream));
    VGTKB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *status = (int64_t)v;
    return VGTKB_OK;
}
