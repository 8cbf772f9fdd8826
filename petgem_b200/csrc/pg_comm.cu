// Peer-memory collectives of the multi-GPU Krylov iteration (one process per GPU, NVLink / NVSwitch).
//
// Reference: the reference leaves both collectives of an iteration to PETSc -- the VecScatter inside
// MatMult and the MPI_Allreduce inside VecDot/VecNorm of KSP.solve (solver.py:584-590), reached through
// createParallelMatrix/createParallelVector (parallel.py:150-203).  On one NVSwitch box neither needs a
// library call: every rank maps the buffers of its peers (CUDA IPC) and
//
//   * the x halo is PUSHED: the owner of the entries writes them straight into the [own | halo] vector of
//     each neighbour over NVLink (pg_comm_push), the neighbour's MatMult is preceded by a flag wait
//     (pg_comm_wait) and followed by an acknowledgement (pg_comm_ack) so that the next push cannot
//     overwrite entries still being read;
//   * the dot products are reduced by ONE single-block kernel per reduction (pg_comm_allreduce): every
//     rank stores its k partial sums into a slot of every peer, raises a flag, waits for the flags of the
//     others and adds the slots in rank order -- bit-identical on every rank, no host involvement.
//
// Everything is a plain kernel with device-resident sequence counters, so the iterations between two host
// checks are captured in a CUDA graph on several GPUs exactly like on one.  Waits are bounded: after
// `timeout` they raise a sticky error (pg_comm_status) instead of hanging the device.
#include <string.h>

#include <algorithm>
#include <cub/cub.cuh>

#include "pg_common.cuh"

namespace pg {
namespace {

constexpr int kMaxRanks = PG_COMM_MAX_RANKS;
constexpr int kMaxChan = PG_COMM_MAX_CHANNELS;
constexpr int kMaxRedDoubles = 2 * PG_COMM_MAX_REDUCE;

// symmetric control block: one per rank, written by the peers
struct Ctrl {
    unsigned long long red_flag[2][kMaxRanks];           // [parity][src]: sequence number of src's slot
    double red_data[2][kMaxRanks][kMaxRedDoubles];       // [parity][src][value]
    unsigned long long data_flag[kMaxChan][kMaxRanks];   // [channel][src]: src has pushed #seq into my buffers
    unsigned long long ack_flag[kMaxChan][kMaxRanks];    // [channel][reader]: reader is done with my push #seq
};

// private counters of this rank
struct Local {
    unsigned long long red_seq;
    unsigned long long push_seq[kMaxChan];
    unsigned long long wait_seq[kMaxChan];
    unsigned int push_done[kMaxChan];
    int error;
};

struct Peers {
    Ctrl *p[kMaxRanks];
};

struct PushArgs {
    double2 *dst[kMaxRanks];
    long long seg[kMaxRanks + 1];
    int dsts[kMaxRanks];  // the destination ranks that get entries, compacted
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(double *p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// spin until *flag >= seq; false (and the sticky error set) after timeout_ns
__device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long seq, Local *loc,
                                          unsigned long long timeout_ns) {
    if (ld_acquire_sys(flag) >= seq) return true;
    if (*reinterpret_cast<volatile int *>(&loc->error)) return false;
    const unsigned long long t0 = now_ns();
    unsigned backoff = 32;
    while (ld_acquire_sys(flag) < seq) {
        __nanosleep(backoff);
        if (backoff < 1024) backoff *= 2;
        if (now_ns() - t0 > timeout_ns) {
            atomicExch(&loc->error, 1);
            return false;
        }
    }
    return true;
}

__global__ void __launch_bounds__(kMaxRedDoubles) allreduce_kernel(Peers pp, Local *loc, int rank, int world, int nd,
                                                                   const double *in, double *out,
                                                                   unsigned long long timeout_ns) {
    __shared__ unsigned long long s_seq;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_seq = loc->red_seq + 1;
        loc->red_seq = s_seq;
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int b = (int)(seq & 1ull);
    if (tid < nd) {
        const double v = in[tid];
        for (int r = 0; r < world; ++r) st_relaxed_sys(&pp.p[r]->red_data[b][rank][tid], v);
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) st_release_sys(&pp.p[tid]->red_flag[b][rank], seq);
    Ctrl *me = pp.p[rank];
    if (tid < world) wait_flag(&me->red_flag[b][tid], seq, loc, timeout_ns);
    __threadfence_system();
    __syncthreads();
    if (tid < nd) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += ld_relaxed_sys(&me->red_data[b][r][tid]);  // rank order on every rank
        if (*reinterpret_cast<volatile int *>(&loc->error)) s = __longlong_as_double(0x7ff8000000000000ll);
        out[tid] = s;
    }
}

// grid (bx, ndst): column j of the grid serves the j-th destination rank that gets entries (a.dsts[j])
__global__ void __launch_bounds__(256) push_kernel(PushArgs a, Peers pp, Local *loc, int rank, int ndst, int chan,
                                                   int k, const double2 *__restrict__ x,
                                                   const int32_t *__restrict__ idx, unsigned long long timeout_ns) {
    const int d = a.dsts[blockIdx.y];
    const long long cnt = a.seg[d + 1] - a.seg[d];
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long *>(&loc->push_seq[chan]) + 1;
    // the reader must be done with the previous push before its halo entries are overwritten
    if (threadIdx.x == 0) wait_flag(&pp.p[rank]->ack_flag[chan][d], seq - 1, loc, timeout_ns);
    __syncthreads();
    const int32_t *id = idx + a.seg[d];
    double2 *dst = a.dst[d];
    const long long total = cnt * k;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long ent = e / k;
        const int r = (int)(e - ent * k);
        dst[e] = x[(long long)__ldg(id + ent) * k + r];
    }
    // bar.sync orders the block's stores before thread 0, whose system-scope fence is cumulative over them
    // (the grid-sync pattern): one fence per block instead of one per thread
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned nblocks = gridDim.x * gridDim.y;
        if (atomicAdd(&loc->push_done[chan], 1u) == nblocks - 1) {  // last block: every entry is on its way
            loc->push_done[chan] = 0;
            loc->push_seq[chan] = seq;
            __threadfence_system();
            for (int j = 0; j < ndst; ++j) st_release_sys(&pp.p[a.dsts[j]]->data_flag[chan][rank], seq);
        }
    }
}

__global__ void __launch_bounds__(32) wait_kernel(Peers pp, Local *loc, int rank, int chan, unsigned from_mask,
                                                  unsigned long long timeout_ns) {
    __shared__ unsigned long long s_seq;
    if (threadIdx.x == 0) {
        s_seq = loc->wait_seq[chan] + 1;
        loc->wait_seq[chan] = s_seq;
    }
    __syncwarp();
    if ((from_mask >> threadIdx.x) & 1u) wait_flag(&pp.p[rank]->data_flag[chan][threadIdx.x], s_seq, loc, timeout_ns);
    __threadfence_system();
}

__global__ void __launch_bounds__(32) ack_kernel(Peers pp, Local *loc, int rank, int chan, unsigned from_mask) {
    const unsigned long long seq = loc->wait_seq[chan];
    if ((from_mask >> threadIdx.x) & 1u) st_release_sys(&pp.p[threadIdx.x]->ack_flag[chan][rank], seq);
}

// ---- halo set-up: which columns of the owned block live on other ranks, and the [own | halo] numbering ----
__global__ void __launch_bounds__(256) flag_outside_kernel(int64_t nnz, const int32_t *__restrict__ col, int32_t lo,
                                                           int32_t hi, int32_t *__restrict__ key) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = col[i];
        key[i] = (c < lo || c >= hi) ? c : 0x7fffffff;  // owned columns sort to the end
    }
}

__global__ void __launch_bounds__(256) remap_halo_kernel(int64_t nnz, int32_t *__restrict__ col, int32_t lo, int32_t hi,
                                                         const int32_t *__restrict__ ext, int32_t n_ext) {
    const int32_t n = hi - lo;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = col[i];
        if (c >= lo && c < hi) {
            col[i] = c - lo;
        } else {  // position in the sorted external list (binary search)
            int32_t a = 0, b = n_ext;
            while (a < b) {
                const int32_t m = (a + b) >> 1;
                if (__ldg(ext + m) < c) a = m + 1; else b = m;
            }
            col[i] = n + a;
        }
    }
}

}  // namespace
}  // namespace pg

using namespace pg;

struct pg_comm {
    int rank, world;
    Peers peers;
    Local *local;
    unsigned long long timeout_ns;
};

extern "C" {

/* ---- raw device memory other processes of the box can map (CUDA IPC) ---- */
int pg_ipc_alloc(int64_t bytes, void **ptr, void *handle_host) {
    PG_REQUIRE(bytes > 0 && ptr && handle_host, PG_EINVAL, "pg_ipc_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == PG_IPC_HANDLE_BYTES, "IPC handle size");
    void *p = nullptr;
    PG_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
    PG_CUDA_OK(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("pg_ipc_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
        return PG_ECUDA;
    }
    memcpy(handle_host, &h, sizeof(h));
    *ptr = p;
    return PG_OK;
}

int pg_ipc_open(const void *handle_host, void **ptr) {
    PG_REQUIRE(handle_host && ptr, PG_EINVAL, "pg_ipc_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    PG_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PG_OK;
}

int pg_ipc_close(void *ptr) {
    if (ptr) PG_CUDA_OK(cudaIpcCloseMemHandle(ptr));
    return PG_OK;
}

int pg_ipc_free(void *ptr) {
    if (ptr) PG_CUDA_OK(cudaFree(ptr));
    return PG_OK;
}

/* ---- halo set-up on the device (what PETSc's MatSetUpMultiply / VecScatterCreate do for an MPIAIJ matrix) ---- */
int pg_halo_columns(int64_t nnz, const int32_t *colidx, int64_t row_begin, int64_t row_end, int32_t *ext,
                    int64_t *n_ext_host, void *stream) {
    PG_REQUIRE(nnz >= 0 && n_ext_host && (nnz == 0 || (colidx && ext)) && row_end >= row_begin, PG_EINVAL,
               "pg_halo_columns: bad argument");
    PG_REQUIRE(nnz < (int64_t)1 << 31, PG_ERANGE, "pg_halo_columns: more than 2^31 nonzeros in one block");
    *n_ext_host = 0;
    if (nnz == 0) return PG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *key = nullptr, *sorted = nullptr, *uniq = nullptr, *count = nullptr;
    void *tmp = nullptr;
    size_t b1 = 0, b2 = 0;
    int rc = PG_OK;
    auto cleanup = [&]() {
        cudaFree(key);
        cudaFree(sorted);
        cudaFree(uniq);
        cudaFree(count);
        cudaFree(tmp);
    };
#define PG_TRYC(expr)                                                                  \
    do {                                                                               \
        cudaError_t e_ = (expr);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            set_error("pg_halo_columns: %s -> %s", #expr, cudaGetErrorString(e_));     \
            cleanup();                                                                 \
            return e_ == cudaErrorMemoryAllocation ? PG_ENOMEM : PG_ECUDA;             \
        }                                                                              \
    } while (0)
    PG_TRYC(cudaMalloc((void **)&key, (size_t)nnz * 4));
    PG_TRYC(cudaMalloc((void **)&sorted, (size_t)nnz * 4));
    PG_TRYC(cudaMalloc((void **)&uniq, (size_t)nnz * 4));
    PG_TRYC(cudaMalloc((void **)&count, 4));
    flag_outside_kernel<<<(unsigned)std::min<int64_t>((nnz + 255) / 256, (int64_t)kNumSMs * 16), 256, 0, st>>>(
        nnz, colidx, (int32_t)row_begin, (int32_t)row_end, key);
    PG_TRYC(cudaGetLastError());
    PG_TRYC(cub::DeviceRadixSort::SortKeys(nullptr, b1, key, sorted, (int)nnz, 0, 32, st));
    PG_TRYC(cub::DeviceSelect::Unique(nullptr, b2, sorted, uniq, count, (int)nnz, st));
    PG_TRYC(cudaMalloc(&tmp, std::max(b1, b2)));
    PG_TRYC(cub::DeviceRadixSort::SortKeys(tmp, b1, key, sorted, (int)nnz, 0, 32, st));
    PG_TRYC(cub::DeviceSelect::Unique(tmp, b2, sorted, uniq, count, (int)nnz, st));
    int32_t nu = 0, last = 0;
    PG_TRYC(cudaMemcpyAsync(&nu, count, 4, cudaMemcpyDeviceToHost, st));
    PG_TRYC(cudaStreamSynchronize(st));
    if (nu > 0) {
        PG_TRYC(cudaMemcpyAsync(&last, uniq + nu - 1, 4, cudaMemcpyDeviceToHost, st));
        PG_TRYC(cudaStreamSynchronize(st));
        if (last == 0x7fffffff) --nu;  // the marker of the owned columns
    }
    if (nu > 0) PG_TRYC(cudaMemcpyAsync(ext, uniq, (size_t)nu * 4, cudaMemcpyDeviceToDevice, st));
    PG_TRYC(cudaStreamSynchronize(st));
#undef PG_TRYC
    *n_ext_host = nu;
    cleanup();
    return rc;
}

int pg_halo_remap(int64_t nnz, int32_t *colidx, int64_t row_begin, int64_t row_end, const int32_t *ext, int64_t n_ext,
                  void *stream) {
    PG_REQUIRE(nnz >= 0 && (nnz == 0 || colidx) && (n_ext == 0 || ext) && row_end >= row_begin, PG_EINVAL,
               "pg_halo_remap: bad argument");
    if (nnz == 0) return PG_OK;
    remap_halo_kernel<<<(unsigned)std::min<int64_t>((nnz + 255) / 256, (int64_t)kNumSMs * 16), 256, 0,
                        (cudaStream_t)stream>>>(nnz, colidx, (int32_t)row_begin, (int32_t)row_end, ext, (int32_t)n_ext);
    PG_LAUNCH_OK();
    return PG_OK;
}

int64_t pg_comm_ctrl_bytes(void) { return (int64_t)sizeof(Ctrl); }

int pg_comm_create(int rank, int world, void *const *ctrl_host, double timeout_s, pg_comm **comm) {
    PG_REQUIRE(comm && ctrl_host && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, PG_EINVAL,
               "pg_comm_create: rank %d of %d (at most %d ranks)", rank, world, kMaxRanks);
    pg_comm *c = new pg_comm();
    c->rank = rank;
    c->world = world;
    for (int r = 0; r < kMaxRanks; ++r) c->peers.p[r] = r < world ? static_cast<Ctrl *>(ctrl_host[r]) : nullptr;
    c->timeout_ns = (unsigned long long)((timeout_s > 0.0 ? timeout_s : 20.0) * 1e9);
    cudaError_t e = cudaMalloc(&c->local, sizeof(Local));
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, sizeof(Local));
    if (e != cudaSuccess) {
        set_error("pg_comm_create: %s", cudaGetErrorString(e));
        delete c;
        return PG_ECUDA;
    }
    *comm = c;
    return PG_OK;
}

void pg_comm_destroy(pg_comm *c) {
    if (!c) return;
    cudaFree(c->local);
    delete c;
}

int pg_comm_allreduce(pg_comm *c, int k, const double *in, double *out, void *stream) {
    PG_REQUIRE(c && in && out && k >= 1 && k <= PG_COMM_MAX_REDUCE, PG_EINVAL,
               "pg_comm_allreduce: k = %d complex scalars (1..%d)", k, PG_COMM_MAX_REDUCE);
    allreduce_kernel<<<1, kMaxRedDoubles, 0, (cudaStream_t)stream>>>(c->peers, c->local, c->rank, c->world, 2 * k, in,
                                                                     out, c->timeout_ns);
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_comm_push(pg_comm *c, int chan, int k, const double *x, const int32_t *send_idx, const int64_t *seg_host,
                 void *const *dst_host, void *stream) {
    PG_REQUIRE(c && x && seg_host && dst_host && chan >= 0 && chan < kMaxChan && k >= 1, PG_EINVAL,
               "pg_comm_push: bad argument");
    PushArgs a;
    long long most = 0;
    for (int r = 0; r < kMaxRanks; ++r) {
        a.dst[r] = r < c->world ? static_cast<double2 *>(dst_host[r]) : nullptr;
        a.seg[r] = seg_host[std::min(r, c->world)];
    }
    a.seg[kMaxRanks] = seg_host[c->world];
    int ndst = 0;
    for (int r = 0; r < kMaxRanks; ++r) a.dsts[r] = 0;
    for (int r = 0; r < c->world; ++r) {
        const long long cnt = seg_host[r + 1] - seg_host[r];
        PG_REQUIRE(cnt >= 0 && (cnt == 0 || (dst_host[r] && send_idx)), PG_EINVAL, "pg_comm_push: segment %d", r);
        most = std::max(most, cnt);
        if (cnt > 0) a.dsts[ndst++] = r;
    }
    if (most == 0) return PG_OK;  // nothing to send to anybody: no flags either (the readers expect none)
    const unsigned bx = (unsigned)std::min<long long>((most * k + 1023) / 1024, 64);  // >= 4 entries per thread
    push_kernel<<<dim3(bx, ndst), 256, 0, (cudaStream_t)stream>>>(a, c->peers, c->local, c->rank, ndst, chan, k,
                                                                  reinterpret_cast<const double2 *>(x), send_idx,
                                                                  c->timeout_ns);
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_comm_wait(pg_comm *c, int chan, uint32_t from_mask, void *stream) {
    PG_REQUIRE(c && chan >= 0 && chan < kMaxChan, PG_EINVAL, "pg_comm_wait: bad argument");
    if (!from_mask) return PG_OK;
    wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c->peers, c->local, c->rank, chan, from_mask, c->timeout_ns);
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_comm_ack(pg_comm *c, int chan, uint32_t from_mask, void *stream) {
    PG_REQUIRE(c && chan >= 0 && chan < kMaxChan, PG_EINVAL, "pg_comm_ack: bad argument");
    if (!from_mask) return PG_OK;
    ack_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c->peers, c->local, c->rank, chan, from_mask);
    PG_LAUNCH_OK();
    return PG_OK;
}

int pg_comm_status(pg_comm *c, void *stream) {
    PG_REQUIRE(c, PG_EINVAL, "pg_comm_status: null");
    int err = 0;
    PG_CUDA_OK(cudaMemcpyAsync(&err, &c->local->error, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PG_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    if (err) {
        set_error("pg_comm: a peer did not answer within %.0f s (rank %d of %d)", c->timeout_ns * 1e-9, c->rank, c->world);
        return PG_ETIMEDOUT;
    }
    return PG_OK;
}

}  // extern "C"
