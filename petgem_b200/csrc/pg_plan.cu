// Symbolic phase: entity incidence lists, per-entity column lists, CSR pattern.
// Replaces the pattern decisions PETSc makes inside MatSetValues/MatAssembly
// (solver.py:188-235) and the numbering of hvfem.py:15-98; runs once per mesh/order.
#include <cub/cub.cuh>
#include <new>

#include "pg_plan.cuh"

namespace pg {

constexpr int kCandCap = 704;       // 64 incident elements x 11 slots per entity
constexpr int kSymWarps = 4;

__global__ void incidence_keys_kernel(int64_t T, int nslots, const int32_t *__restrict__ elemsE,
                                      const int32_t *__restrict__ elemsF, int64_t nE, int64_t nF,
                                      int32_t *__restrict__ keys, int32_t *__restrict__ vals,
                                      int32_t *__restrict__ counts) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= T * nslots) return;
    int64_t t = i / nslots;
    int s = (int)(i - t * nslots);
    int32_t g = (int32_t)global_entity(elemsE, elemsF, nE, nF, t, s);
    keys[i] = g;
    vals[i] = (int32_t)i;
    atomicAdd(counts + g, 1);  // integer histogram: order-independent result
}

__global__ void iota_kernel(int64_t n, int32_t *out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)i;
}

__global__ void invert_order_kernel(int64_t n, const int32_t *__restrict__ order, int32_t *__restrict__ inv,
                                    int64_t *__restrict__ rows, int64_t nE, int64_t nF, int p, int *bad) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= n) return;
    int32_t g = order[b];
    if (g < 0 || g >= n) {
        atomicExch(bad, 1);
        return;
    }
    if (atomicExch(inv + g, (int32_t)b) != -1) atomicExch(bad, 1);  // not a permutation
    rows[b] = rows_of_entity(g, nE, nF, p);
}

// first incidence value of each entity (element-major locality key)
__global__ void first_incidence_kernel(int64_t nEnt, int nslots, const int32_t *__restrict__ inc_ptr,
                                       const int32_t *__restrict__ sorted_vals,
                                       const int32_t *__restrict__ elem_rank, int32_t *__restrict__ key) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= nEnt) return;
    int32_t a = inc_ptr[g], b = inc_ptr[g + 1];
    int32_t k = 2147483647;
    if (b > a) {
        if (!elem_rank) {
            k = sorted_vals[a];  // incidences are sorted by element: the first one is the smallest
        } else {                 // entity key = earliest incident element in the caller's traversal
            for (int32_t i = a; i < b; ++i) {
                const int32_t v = sorted_vals[i], t = v / nslots;
                k = min(k, elem_rank[t] * nslots + (v - t * nslots));
            }
        }
    }
    key[g] = k;
}

// PASS 0: count unique column entities and the row length; PASS 1: fill lists and records.
template <int PASS>
__global__ void __launch_bounds__(kSymWarps * 32)
    colent_kernel(int64_t nEnt, int p, int nslots, const int32_t *__restrict__ elemsE,
                  const int32_t *__restrict__ elemsF, int64_t nE, int64_t nF, const int32_t *__restrict__ inc_ptr,
                  const int32_t *__restrict__ sorted_vals, const int32_t *__restrict__ blk_of_ent,
                  int32_t *__restrict__ ncol, int32_t *__restrict__ rowlen, const int64_t *__restrict__ colent_ptr,
                  int32_t *__restrict__ colent, int32_t *__restrict__ selfpos, IncRecord *__restrict__ rec,
                  int *__restrict__ overflow) {
    __shared__ int32_t s_blk[kSymWarps][kCandCap];
    __shared__ int32_t s_ent[kSymWarps][kCandCap];
    __shared__ uint16_t s_rank[kSymWarps][kCandCap];
    __shared__ uint16_t s_pos[kSymWarps][kCandCap];   // by rank
    __shared__ uint8_t s_rows[kSymWarps][kCandCap];   // by rank
    __shared__ uint8_t s_first[kSymWarps][kCandCap];

    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t warp = blockIdx.x * (int64_t)kSymWarps + w;
    const int64_t nwarps = (int64_t)gridDim.x * kSymWarps;
    int32_t *blk = s_blk[w], *ent = s_ent[w];
    uint16_t *rank = s_rank[w], *pos = s_pos[w];
    uint8_t *rws = s_rows[w], *first = s_first[w];

    for (int64_t g = warp; g < nEnt; g += nwarps) {
        const int32_t i0 = inc_ptr[g], m = inc_ptr[g + 1] - i0;
        const int M = m * nslots;
        if (M > kCandCap) {
            if (lane == 0) atomicExch(overflow, 1);
            continue;
        }
        for (int i = lane; i < M; i += 32) {
            int a = i / nslots, s = i - a * nslots;
            int64_t t = sorted_vals[i0 + a] / nslots;
            int32_t e = (int32_t)global_entity(elemsE, elemsF, nE, nF, t, s);
            ent[i] = e;
            blk[i] = blk_of_ent[e];
        }
        __syncwarp();
        int nfirst_local = 0;
        for (int i = lane; i < M; i += 32) {
            int32_t key = blk[i];
            bool f = true;
            for (int j = 0; j < i; ++j)
                if (blk[j] == key) {
                    f = false;
                    break;
                }
            first[i] = f;
            nfirst_local += f;
        }
        __syncwarp();
        for (int i = lane; i < M; i += 32) {
            int32_t key = blk[i];
            int r = 0;
            for (int j = 0; j < M; ++j) r += (first[j] && blk[j] < key);
            rank[i] = (uint16_t)r;
            if (first[i]) rws[r] = (uint8_t)rows_of_entity(ent[i], nE, nF, p);
        }
        int nu = nfirst_local;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nu += __shfl_xor_sync(0xffffffffu, nu, o);
        __syncwarp();
        // exclusive prefix of rows over ranks -> positions
        int carry = 0;
        for (int base = 0; base < nu; base += 32) {
            int r = base + lane;
            int v = (r < nu) ? rws[r] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (r < nu) pos[r] = (uint16_t)(carry + incl - v);
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        if (PASS == 0) {
            if (lane == 0) {
                ncol[g] = nu;
                rowlen[g] = carry;
            }
        } else {
            const int64_t c0 = colent_ptr[g];
            for (int i = lane; i < M; i += 32) {
                if (first[i]) colent[c0 + rank[i]] = ent[i];
                if (ent[i] == (int32_t)g) selfpos[g] = pos[rank[i]];  // same value from every writer
            }
            for (int a = lane; a < m; a += 32) {
                IncRecord r;
                int32_t v = sorted_vals[i0 + a];
                r.elem = v / nslots;
                r.slot = (uint8_t)(v - r.elem * nslots);
                r.pad0 = 0;
                r.bdmask = 0;
                uint16_t fm = 0;
#pragma unroll
                for (int s = 0; s < PG_SLOTS; ++s) {
                    r.slotpos[s] = (s < nslots) ? pos[rank[a * nslots + s]] : 0;
                    if (s < nslots && first[a * nslots + s]) fm |= (uint16_t)(1u << s);
                }
                r.firstmask = fm;
                rec[i0 + a] = r;
            }
        }
        __syncwarp();
    }
}

__global__ void colstart_kernel(int64_t n, const int32_t *__restrict__ colent,
                                const int32_t *__restrict__ blk_of_ent, const int64_t *__restrict__ row_base,
                                int32_t *__restrict__ colstart) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) colstart[i] = (int32_t)row_base[blk_of_ent[colent[i]]];
}

__global__ void block_lengths_kernel(int64_t b0, int64_t b1, int n, const int32_t *__restrict__ ent_order,
                                     const int32_t *__restrict__ rowlen, const int32_t *__restrict__ inc_ptr,
                                     const int64_t *__restrict__ row_base, int64_t *__restrict__ vlen,
                                     int64_t *__restrict__ contrib) {
    int64_t b = b0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= b1) return;
    int32_t g = ent_order[b];
    int64_t r = row_base[b + 1] - row_base[b];
    vlen[b - b0] = r * rowlen[g];
    contrib[b - b0] = r * (int64_t)(inc_ptr[g + 1] - inc_ptr[g]) * n;
}

// element range touched by the owned blocks (records of an entity are sorted by element)
__global__ void elem_range_kernel(int64_t b0, int64_t b1, const int32_t *__restrict__ ent_order,
                                  const int32_t *__restrict__ inc_ptr, const IncRecord *__restrict__ rec,
                                  int *__restrict__ mm) {
    int64_t b = b0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= b1) return;
    const int32_t g = ent_order[b];
    const int32_t a = inc_ptr[g], e = inc_ptr[g + 1];
    if (e > a) {
        atomicMin(mm, rec[a].elem);
        atomicMax(mm + 1, rec[e - 1].elem);
    }
}

// block range of a row range; res = {b0, b1, aligned_begin, aligned_end}
__global__ void find_blocks_kernel(int64_t nEnt, const int64_t *__restrict__ row_base, int64_t row_begin,
                                   int64_t row_end, int64_t *res) {
    auto lower = [&](int64_t row) {
        int64_t lo = 0, hi = nEnt;  // first b with row_base[b] >= row
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (row_base[mid] < row) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    int64_t a = lower(row_begin), b = lower(row_end);
    res[0] = a;
    res[1] = b;
    res[2] = row_base[a];
    res[3] = row_base[b];
}

// CSR of the owned rows, warp per block
__global__ void __launch_bounds__(256)
    csr_kernel(int64_t b0, int64_t b1, int p, int64_t nE, int64_t nF, const int32_t *__restrict__ ent_order,
               const int32_t *__restrict__ blk_of_ent, const int64_t *__restrict__ row_base,
               const int64_t *__restrict__ colent_ptr, const int32_t *__restrict__ colent,
               const int32_t *__restrict__ rowlen, const int64_t *__restrict__ valoff, int64_t row_begin,
               int64_t *__restrict__ rowptr, int32_t *__restrict__ colidx) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = b0 + warp; b < b1; b += nwarps) {
        const int32_t g = ent_order[b];
        const int r = (int)(row_base[b + 1] - row_base[b]);
        const int L = rowlen[g];
        const int64_t v0 = valoff[b - b0];
        const int64_t row0 = row_base[b] - row_begin;
        for (int d = lane; d < r; d += 32) rowptr[row0 + d] = v0 + (int64_t)d * L;
        const int64_t c0 = colent_ptr[g];
        const int nc = (int)(colent_ptr[g + 1] - c0);
        int posbase = 0;
        for (int base = 0; base < nc; base += 32) {
            int i = base + lane;
            int32_t e = (i < nc) ? colent[c0 + i] : 0;
            int rr = (i < nc) ? rows_of_entity(e, nE, nF, p) : 0;
            int incl = rr;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (i < nc) {
                int pos = posbase + incl - rr;
                int32_t col0 = (int32_t)row_base[blk_of_ent[e]];
                for (int d = 0; d < r; ++d)
                    for (int dd = 0; dd < rr; ++dd) colidx[v0 + (int64_t)d * L + pos + dd] = col0 + dd;
            }
            posbase += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

__global__ void dof_perm_kernel(int64_t nEnt, int p, int64_t nE, int64_t nF, int64_t T,
                                const int32_t *__restrict__ blk_of_ent, const int64_t *__restrict__ row_base,
                                int32_t *__restrict__ perm) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= nEnt) return;
    int r = rows_of_entity(g, nE, nF, p);
    int64_t ref0;
    if (g < nE) ref0 = g * ndof_edge(p);
    else if (g < nE + nF) ref0 = nE * ndof_edge(p) + (g - nE) * ndof_face(p);
    else ref0 = nE * ndof_edge(p) + nF * ndof_face(p) + (g - nE - nF) * ndof_volume(p);
    int64_t new0 = row_base[blk_of_ent[g]];
    for (int d = 0; d < r; ++d) perm[ref0 + d] = (int32_t)(new0 + d);
}

__global__ void set_bdmask_kernel(int64_t nInc, int nslots, const int32_t *__restrict__ elemsE,
                                  const int32_t *__restrict__ elemsF, int64_t nE, int64_t nF,
                                  const uint8_t *__restrict__ bd_entity, IncRecord *__restrict__ rec) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nInc) return;
    int64_t t = rec[i].elem;
    uint16_t m = 0;
    for (int s = 0; s < nslots; ++s)
        if (bd_entity[global_entity(elemsE, elemsF, nE, nF, t, s)]) m |= (uint16_t)(1u << s);
    rec[i].bdmask = m;
}

constexpr int kProcWindow = 16384;  // entities per window of the processing order

__global__ void hdr_keys_kernel(int64_t b0, int64_t nb, const int32_t *__restrict__ ent_order,
                                const int32_t *__restrict__ inc_ptr, const uint8_t *__restrict__ bd_entity,
                                int32_t *__restrict__ keys, int32_t *__restrict__ vals) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const int32_t g = ent_order[b0 + i];
    int m = inc_ptr[g + 1] - inc_ptr[g];
    if (m > 255) m = 255;
    if (bd_entity && bd_entity[g]) m = 0;
    keys[i] = (int32_t)(((i / kProcWindow) << 8) | m);
    vals[i] = (int32_t)i;
}

__global__ void hdr_fill_kernel(int64_t b0, int64_t nb, const int32_t *__restrict__ order,
                                const int32_t *__restrict__ ent_order, const int64_t *__restrict__ row_base,
                                const int32_t *__restrict__ inc_ptr, const int32_t *__restrict__ rowlen,
                                const int32_t *__restrict__ selfpos, const int64_t *__restrict__ valoff,
                                const int64_t *__restrict__ colent_ptr, const uint8_t *__restrict__ bd_entity,
                                EntHdr *__restrict__ hdr) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const int64_t bl = order[i];
    const int32_t g = ent_order[b0 + bl];
    EntHdr h;
    h.valoff = valoff[bl];
    h.inc0 = inc_ptr[g];
    h.ent = g;
    h.m = (uint16_t)(inc_ptr[g + 1] - inc_ptr[g]);
    h.L = (uint16_t)rowlen[g];
    h.selfpos = (uint16_t)selfpos[g];
    h.bd = (bd_entity && bd_entity[g]) ? 1 : 0;
    h.rows = (uint8_t)(row_base[b0 + bl + 1] - row_base[b0 + bl]);
    h.row = (int32_t)(row_base[b0 + bl] - row_base[b0]);
    h.cbase = (int32_t)colent_ptr[g];
    hdr[i] = h;
}

static int bits_for(int64_t n) {
    int b = 1;
    while ((1LL << b) < n) ++b;
    return b;
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <typename T> T *as() { return static_cast<T *>(p); }
};

// sorted incidence lists: inc_ptr [nEnt+1], sorted_vals [nInc]
static int build_incidence(int64_t T, int p, const int32_t *elemsE, const int32_t *elemsF, int64_t nE, int64_t nF,
                           int64_t nEnt, int32_t *inc_ptr, int32_t *sorted_vals, cudaStream_t st) {
    const int nslots = nslots_of(p);
    const int64_t nInc = T * nslots;
    DevBuf keys, vals, keys_out, tmp;
    PG_CUDA_OK(cudaMalloc(&keys.p, nInc * 4));
    PG_CUDA_OK(cudaMalloc(&vals.p, nInc * 4));
    PG_CUDA_OK(cudaMalloc(&keys_out.p, nInc * 4));
    PG_CUDA_OK(cudaMemsetAsync(inc_ptr, 0, (nEnt + 1) * 4, st));
    incidence_keys_kernel<<<(unsigned)((nInc + 255) / 256), 256, 0, st>>>(T, nslots, elemsE, elemsF, nE, nF,
                                                                         keys.as<int32_t>(), vals.as<int32_t>(),
                                                                         inc_ptr);
    PG_LAUNCH_OK();
    size_t bytes = 0;
    PG_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.as<int32_t>(), keys_out.as<int32_t>(),
                                               vals.as<int32_t>(), sorted_vals, nInc, 0, bits_for(nEnt), st));
    PG_CUDA_OK(cudaMalloc(&tmp.p, bytes));
    PG_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys.as<int32_t>(), keys_out.as<int32_t>(),
                                               vals.as<int32_t>(), sorted_vals, nInc, 0, bits_for(nEnt), st));
    DevBuf tmp2;
    bytes = 0;
    PG_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, inc_ptr, inc_ptr, nEnt + 1, st));
    PG_CUDA_OK(cudaMalloc(&tmp2.p, bytes));
    PG_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp2.p, bytes, inc_ptr, inc_ptr, nEnt + 1, st));
    PG_CUDA_OK(cudaStreamSynchronize(st));
    return PG_OK;
}

// headers of the owned entities in processing order (rebuilt when the Dirichlet set changes)
static int build_headers(pg_plan *pl, cudaStream_t st) {
    const int64_t nb = pl->b1 - pl->b0;
    if (nb <= 0) return PG_OK;
    if (!pl->hdr) PG_CUDA_OK(cudaMalloc((void **)&pl->hdr, nb * sizeof(EntHdr)));
    DevBuf keys, vals, keys_out, order, tmp;
    PG_CUDA_OK(cudaMalloc(&keys.p, nb * 4));
    PG_CUDA_OK(cudaMalloc(&vals.p, nb * 4));
    PG_CUDA_OK(cudaMalloc(&keys_out.p, nb * 4));
    PG_CUDA_OK(cudaMalloc(&order.p, nb * 4));
    const unsigned grid = (unsigned)((nb + 255) / 256);
    hdr_keys_kernel<<<grid, 256, 0, st>>>(pl->b0, nb, pl->ent_order, pl->inc_ptr, pl->bd_entity,
                                          keys.as<int32_t>(), vals.as<int32_t>());
    PG_LAUNCH_OK();
    const int bits = 8 + bits_for(nb / kProcWindow + 2);
    size_t bytes = 0;
    PG_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.as<int32_t>(), keys_out.as<int32_t>(),
                                               vals.as<int32_t>(), order.as<int32_t>(), nb, 0, bits, st));
    PG_CUDA_OK(cudaMalloc(&tmp.p, bytes));
    PG_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys.as<int32_t>(), keys_out.as<int32_t>(),
                                               vals.as<int32_t>(), order.as<int32_t>(), nb, 0, bits, st));
    hdr_fill_kernel<<<grid, 256, 0, st>>>(pl->b0, nb, order.as<int32_t>(), pl->ent_order, pl->row_base, pl->inc_ptr,
                                          pl->rowlen, pl->selfpos, pl->valoff, pl->colent_ptr, pl->bd_entity, pl->hdr);
    PG_LAUNCH_OK();
    PG_CUDA_OK(cudaStreamSynchronize(st));
    return PG_OK;
}

static int check_sizes(int64_t T, int p, int64_t nE, int64_t nF, int64_t &nEnt, int64_t &N) {
    PG_REQUIRE(p >= 1 && p <= PG_MAX_ORDER, PG_EINVAL, "polynomial order %d outside 1..%d", p, PG_MAX_ORDER);
    PG_REQUIRE(T > 0 && nE > 0 && nF > 0, PG_EINVAL, "plan: empty mesh (T=%lld nE=%lld nF=%lld)", (long long)T,
               (long long)nE, (long long)nF);
    nEnt = nE + (p >= 2 ? nF : 0) + (p >= 3 ? T : 0);
    N = nE * ndof_edge(p) + nF * ndof_face(p) + T * ndof_volume(p);
    PG_REQUIRE(N < 2147483647LL && T * (int64_t)PG_SLOTS < 2147483647LL, PG_ERANGE,
               "plan: %lld dofs / %lld elements exceed the int32 index range", (long long)N, (long long)T);
    return PG_OK;
}

}  // namespace pg

using namespace pg;

extern "C" {

void pg_plan_destroy(pg_plan *pl) {
    if (!pl) return;
    cudaFree(pl->ent_order);
    cudaFree(pl->blk_of_ent);
    cudaFree(pl->row_base);
    cudaFree(pl->inc_ptr);
    cudaFree(pl->rec);
    cudaFree(pl->colent_ptr);
    cudaFree(pl->colent);
    cudaFree(pl->colstart);
    cudaFree(pl->rowlen);
    cudaFree(pl->selfpos);
    cudaFree(pl->valoff);
    cudaFree(pl->hdr);
    cudaFree(pl->itable);
    cudaFree(pl->bd_entity);
    delete pl;
}

int pg_plan_locality_order(int64_t T, int p, const int32_t *elemsE, const int32_t *elemsF, int64_t nE, int64_t nF,
                           int32_t *ent_order_host, void *stream) {
    return pg_plan_ranked_order(T, p, elemsE, elemsF, nE, nF, nullptr, ent_order_host, stream);
}

int pg_plan_ranked_order(int64_t T, int p, const int32_t *elemsE, const int32_t *elemsF, int64_t nE, int64_t nF,
                         const int32_t *elem_rank, int32_t *ent_order_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int64_t nEnt, N;
    int rc = check_sizes(T, p, nE, nF, nEnt, N);
    if (rc) return rc;
    PG_REQUIRE(elemsE && elemsF && ent_order_host, PG_EINVAL, "pg_plan_locality_order: null pointer");
    const int64_t nInc = T * nslots_of(p);
    DevBuf inc_ptr, sorted_vals, key, key_out, ids, ids_out, tmp;
    PG_CUDA_OK(cudaMalloc(&inc_ptr.p, (nEnt + 1) * 4));
    PG_CUDA_OK(cudaMalloc(&sorted_vals.p, nInc * 4));
    rc = build_incidence(T, p, elemsE, elemsF, nE, nF, nEnt, inc_ptr.as<int32_t>(), sorted_vals.as<int32_t>(), st);
    if (rc) return rc;
    PG_CUDA_OK(cudaMalloc(&key.p, nEnt * 4));
    PG_CUDA_OK(cudaMalloc(&key_out.p, nEnt * 4));
    PG_CUDA_OK(cudaMalloc(&ids.p, nEnt * 4));
    PG_CUDA_OK(cudaMalloc(&ids_out.p, nEnt * 4));
    const unsigned gb = (unsigned)((nEnt + 255) / 256);
    first_incidence_kernel<<<gb, 256, 0, st>>>(nEnt, nslots_of(p), inc_ptr.as<int32_t>(), sorted_vals.as<int32_t>(),
                                               elem_rank, key.as<int32_t>());
    PG_LAUNCH_OK();
    iota_kernel<<<gb, 256, 0, st>>>(nEnt, ids.as<int32_t>());
    PG_LAUNCH_OK();
    size_t bytes = 0;
    PG_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key.as<int32_t>(), key_out.as<int32_t>(),
                                               ids.as<int32_t>(), ids_out.as<int32_t>(), nEnt, 0, 31, st));
    PG_CUDA_OK(cudaMalloc(&tmp.p, bytes));
    PG_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, key.as<int32_t>(), key_out.as<int32_t>(),
                                               ids.as<int32_t>(), ids_out.as<int32_t>(), nEnt, 0, 31, st));
    PG_CUDA_OK(cudaMemcpyAsync(ent_order_host, ids_out.p, nEnt * 4, cudaMemcpyDeviceToHost, st));
    PG_CUDA_OK(cudaStreamSynchronize(st));
    return PG_OK;
}

int pg_plan_create(int64_t T, int p, const int32_t *elemsE, const int32_t *elemsF, int64_t nE, int64_t nF,
                   const int32_t *ent_order_host, int64_t row_begin, int64_t row_end, pg_plan **out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_REQUIRE(out, PG_EINVAL, "pg_plan_create: null output");
    *out = nullptr;
    int64_t nEnt, N;
    int rc = check_sizes(T, p, nE, nF, nEnt, N);
    if (rc) return rc;
    PG_REQUIRE(elemsE && elemsF, PG_EINVAL, "pg_plan_create: null connectivity");
    if (row_end < 0) row_end = N;
    PG_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= N, PG_EINVAL,
               "pg_plan_create: bad row range [%lld,%lld) of %lld", (long long)row_begin, (long long)row_end,
               (long long)N);

    pg_plan *pl = new (std::nothrow) pg_plan();
    PG_REQUIRE(pl, PG_ENOMEM, "pg_plan_create: host allocation failed");
    struct Guard {
        pg_plan *p;
        ~Guard() { if (p) pg_plan_destroy(p); }
    } guard{pl};

    pl->T = T, pl->p = p, pl->n = ndof_element(p), pl->nslots = nslots_of(p);
    pl->nE = nE, pl->nF = nF, pl->nEnt = nEnt, pl->N = N;
    pl->nInc = T * pl->nslots;
    const unsigned gEnt = (unsigned)((nEnt + 255) / 256);

    // --- incidence lists -------------------------------------------------------
    DevBuf sorted_vals;
    PG_CUDA_OK(cudaMalloc((void **)&pl->inc_ptr, (nEnt + 1) * 4));
    PG_CUDA_OK(cudaMalloc(&sorted_vals.p, pl->nInc * 4));
    rc = build_incidence(T, p, elemsE, elemsF, nE, nF, nEnt, pl->inc_ptr, sorted_vals.as<int32_t>(), st);
    if (rc) return rc;

    // --- entity order and row bases ----------------------------------------------
    PG_CUDA_OK(cudaMalloc((void **)&pl->ent_order, nEnt * 4));
    PG_CUDA_OK(cudaMalloc((void **)&pl->blk_of_ent, nEnt * 4));
    PG_CUDA_OK(cudaMalloc((void **)&pl->row_base, (nEnt + 1) * 8));
    if (ent_order_host)
        PG_CUDA_OK(cudaMemcpyAsync(pl->ent_order, ent_order_host, nEnt * 4, cudaMemcpyHostToDevice, st));
    else {
        iota_kernel<<<gEnt, 256, 0, st>>>(nEnt, pl->ent_order);
        PG_LAUNCH_OK();
    }
    DevBuf flag, tmp;
    PG_CUDA_OK(cudaMalloc(&flag.p, 64));
    PG_CUDA_OK(cudaMemsetAsync(flag.p, 0, 64, st));
    PG_CUDA_OK(cudaMemsetAsync(pl->blk_of_ent, 0xff, nEnt * 4, st));
    PG_CUDA_OK(cudaMemsetAsync(pl->row_base, 0, (nEnt + 1) * 8, st));
    invert_order_kernel<<<gEnt, 256, 0, st>>>(nEnt, pl->ent_order, pl->blk_of_ent, pl->row_base, nE, nF, p,
                                             flag.as<int>());
    PG_LAUNCH_OK();
    {
        size_t bytes = 0;
        PG_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, pl->row_base, pl->row_base, nEnt + 1, st));
        PG_CUDA_OK(cudaMalloc(&tmp.p, bytes));
        PG_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, pl->row_base, pl->row_base, nEnt + 1, st));
    }
    int hflag[2] = {0, 0};
    PG_CUDA_OK(cudaMemcpyAsync(hflag, flag.p, 4, cudaMemcpyDeviceToHost, st));
    PG_CUDA_OK(cudaStreamSynchronize(st));
    PG_REQUIRE(hflag[0] == 0, PG_EINVAL, "pg_plan_create: ent_order is not a permutation of 0..%lld",
               (long long)(nEnt - 1));

    // --- owned block range ---------------------------------------------------------
    {
        DevBuf res;
        int64_t hres[4];
        PG_CUDA_OK(cudaMalloc(&res.p, 32));
        find_blocks_kernel<<<1, 1, 0, st>>>(nEnt, pl->row_base, row_begin, row_end, res.as<int64_t>());
        PG_LAUNCH_OK();
        PG_CUDA_OK(cudaMemcpyAsync(hres, res.p, 32, cudaMemcpyDeviceToHost, st));
        PG_CUDA_OK(cudaStreamSynchronize(st));
        PG_REQUIRE(hres[2] == row_begin && hres[3] == row_end, PG_EINVAL,
                   "pg_plan_create: row range [%lld,%lld) does not fall on entity boundaries (nearest %lld,%lld)",
                   (long long)row_begin, (long long)row_end, (long long)hres[2], (long long)hres[3]);
        pl->b0 = hres[0], pl->b1 = hres[1], pl->row_begin = row_begin, pl->row_end = row_end;
    }

    // --- column-entity lists ----------------------------------------------------------
    DevBuf ncol;
    PG_CUDA_OK(cudaMalloc(&ncol.p, (nEnt + 1) * 4));
    PG_CUDA_OK(cudaMemsetAsync(ncol.p, 0, (nEnt + 1) * 4, st));
    PG_CUDA_OK(cudaMalloc((void **)&pl->rowlen, nEnt * 4));
    PG_CUDA_OK(cudaMalloc((void **)&pl->selfpos, nEnt * 4));
    PG_CUDA_OK(cudaMalloc((void **)&pl->colent_ptr, (nEnt + 1) * 8));
    const unsigned gSym = (unsigned)std::min<int64_t>((nEnt + kSymWarps - 1) / kSymWarps, (int64_t)kNumSMs * 64);
    colent_kernel<0><<<gSym, kSymWarps * 32, 0, st>>>(nEnt, p, pl->nslots, elemsE, elemsF, nE, nF, pl->inc_ptr,
                                                     sorted_vals.as<int32_t>(), pl->blk_of_ent,
                                                     ncol.as<int32_t>(), pl->rowlen, nullptr, nullptr, nullptr,
                                                     nullptr, flag.as<int>() + 1);
    PG_LAUNCH_OK();
    {
        // colent_ptr = exclusive scan of ncol (int32 in, int64 out)
        size_t bytes = 0;
        DevBuf t2;
        PG_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, ncol.as<int32_t>(), pl->colent_ptr, nEnt + 1, st));
        PG_CUDA_OK(cudaMalloc(&t2.p, bytes));
        PG_CUDA_OK(cub::DeviceScan::ExclusiveSum(t2.p, bytes, ncol.as<int32_t>(), pl->colent_ptr, nEnt + 1, st));
        PG_CUDA_OK(cudaStreamSynchronize(st));
    }
    PG_CUDA_OK(cudaMemcpyAsync(hflag, flag.as<int>() + 1, 4, cudaMemcpyDeviceToHost, st));
    int64_t ncolent = 0;
    PG_CUDA_OK(cudaMemcpyAsync(&ncolent, pl->colent_ptr + nEnt, 8, cudaMemcpyDeviceToHost, st));
    PG_CUDA_OK(cudaStreamSynchronize(st));
    PG_REQUIRE(ncolent < 2147483647LL, PG_ERANGE, "pg_plan_create: %lld column entities exceed int32",
               (long long)ncolent);
    PG_REQUIRE(hflag[0] == 0, PG_ERANGE, "pg_plan_create: an entity has more than %d incident elements",
               kCandCap / PG_SLOTS);
    PG_CUDA_OK(cudaMalloc((void **)&pl->colent, std::max<int64_t>(ncolent, 1) * 4));
    PG_CUDA_OK(cudaMalloc((void **)&pl->rec, pl->nInc * sizeof(IncRecord)));
    colent_kernel<1><<<gSym, kSymWarps * 32, 0, st>>>(nEnt, p, pl->nslots, elemsE, elemsF, nE, nF, pl->inc_ptr,
                                                     sorted_vals.as<int32_t>(), pl->blk_of_ent,
                                                     ncol.as<int32_t>(), pl->rowlen, pl->colent_ptr, pl->colent,
                                                     pl->selfpos, pl->rec, flag.as<int>() + 1);
    PG_LAUNCH_OK();

    pl->ncolent = ncolent;
    PG_CUDA_OK(cudaMalloc((void **)&pl->colstart, std::max<int64_t>(ncolent, 1) * 4));
    if (ncolent > 0) {
        colstart_kernel<<<(unsigned)((ncolent + 255) / 256), 256, 0, st>>>(ncolent, pl->colent, pl->blk_of_ent,
                                                                          pl->row_base, pl->colstart);
        PG_LAUNCH_OK();
    }

    // --- value offsets of the owned blocks ---------------------------------------------
    const int64_t nb = pl->b1 - pl->b0;
    PG_CUDA_OK(cudaMalloc((void **)&pl->valoff, (nb + 1) * 8));
    PG_CUDA_OK(cudaMemsetAsync(pl->valoff, 0, (nb + 1) * 8, st));
    if (nb > 0) {
        DevBuf contrib, t3, t4, sum, mx;
        PG_CUDA_OK(cudaMalloc(&contrib.p, nb * 8));
        PG_CUDA_OK(cudaMalloc(&sum.p, 8));
        PG_CUDA_OK(cudaMalloc(&mx.p, 4));
        block_lengths_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(pl->b0, pl->b1, pl->n, pl->ent_order,
                                                                          pl->rowlen, pl->inc_ptr, pl->row_base,
                                                                          pl->valoff, contrib.as<int64_t>());
        PG_LAUNCH_OK();
        size_t bytes = 0;
        PG_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, pl->valoff, pl->valoff, nb + 1, st));
        PG_CUDA_OK(cudaMalloc(&t3.p, bytes));
        PG_CUDA_OK(cub::DeviceScan::ExclusiveSum(t3.p, bytes, pl->valoff, pl->valoff, nb + 1, st));
        bytes = 0;
        PG_CUDA_OK(cub::DeviceReduce::Sum(nullptr, bytes, contrib.as<int64_t>(), sum.as<int64_t>(), nb, st));
        size_t bytes2 = 0;
        PG_CUDA_OK(cub::DeviceReduce::Max(nullptr, bytes2, pl->rowlen, mx.as<int32_t>(), nEnt, st));
        PG_CUDA_OK(cudaMalloc(&t4.p, std::max(bytes, bytes2)));
        PG_CUDA_OK(cub::DeviceReduce::Sum(t4.p, bytes, contrib.as<int64_t>(), sum.as<int64_t>(), nb, st));
        PG_CUDA_OK(cudaMemcpyAsync(&pl->contributions, sum.p, 8, cudaMemcpyDeviceToHost, st));
        PG_CUDA_OK(cub::DeviceReduce::Max(t4.p, bytes2, pl->rowlen, mx.as<int32_t>(), nEnt, st));
        PG_CUDA_OK(cudaMemcpyAsync(&pl->max_rowlen, mx.p, 4, cudaMemcpyDeviceToHost, st));
        PG_CUDA_OK(cudaMemcpyAsync(&pl->nnz, pl->valoff + nb, 8, cudaMemcpyDeviceToHost, st));
        DevBuf mm;
        int hmm[2] = {2147483647, -1};
        PG_CUDA_OK(cudaMalloc(&mm.p, 8));
        PG_CUDA_OK(cudaMemcpyAsync(mm.p, hmm, 8, cudaMemcpyHostToDevice, st));
        elem_range_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(pl->b0, pl->b1, pl->ent_order, pl->inc_ptr,
                                                                       pl->rec, mm.as<int>());
        PG_LAUNCH_OK();
        PG_CUDA_OK(cudaMemcpyAsync(hmm, mm.p, 8, cudaMemcpyDeviceToHost, st));
        PG_CUDA_OK(cudaStreamSynchronize(st));
        pl->elem_begin = hmm[1] >= 0 ? hmm[0] : 0;
        pl->elem_end = hmm[1] >= 0 ? hmm[1] + 1 : 0;
        PG_CUDA_OK(cudaStreamSynchronize(st));
    }
    PG_CUDA_OK(cudaStreamSynchronize(st));
    rc = build_headers(pl, st);
    if (rc) return rc;
    guard.p = nullptr;
    *out = pl;
    return PG_OK;
}

int64_t pg_plan_num_dofs(const pg_plan *pl) { return pl ? pl->N : -1; }
int64_t pg_plan_num_entities(const pg_plan *pl) { return pl ? pl->nEnt : -1; }
int64_t pg_plan_local_rows(const pg_plan *pl) { return pl ? pl->row_end - pl->row_begin : -1; }
int64_t pg_plan_row_begin(const pg_plan *pl) { return pl ? pl->row_begin : -1; }
int64_t pg_plan_nnz(const pg_plan *pl) { return pl ? pl->nnz : -1; }
int64_t pg_plan_contributions(const pg_plan *pl) { return pl ? pl->contributions : -1; }
int pg_plan_max_row_length(const pg_plan *pl) { return pl ? pl->max_rowlen : -1; }
int pg_plan_element_range(const pg_plan *pl, int64_t *t0, int64_t *t1) {
    PG_REQUIRE(pl && t0 && t1, PG_EINVAL, "pg_plan_element_range: null pointer");
    *t0 = pl->elem_begin;
    *t1 = pl->elem_end;
    return PG_OK;
}

int pg_plan_csr(const pg_plan *pl, int64_t *rowptr, int32_t *colidx, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_REQUIRE(pl && rowptr && (colidx || pl->nnz == 0), PG_EINVAL, "pg_plan_csr: null pointer");
    const int64_t nb = pl->b1 - pl->b0;
    const int64_t rows = pl->row_end - pl->row_begin;
    if (nb > 0) {
        unsigned grid = (unsigned)std::min<int64_t>((nb + 7) / 8, (int64_t)kNumSMs * 32);
        csr_kernel<<<grid, 256, 0, st>>>(pl->b0, pl->b1, pl->p, pl->nE, pl->nF, pl->ent_order, pl->blk_of_ent,
                                         pl->row_base, pl->colent_ptr, pl->colent, pl->rowlen, pl->valoff,
                                         pl->row_begin, rowptr, colidx);
        PG_LAUNCH_OK();
    }
    PG_CUDA_OK(cudaMemcpyAsync(rowptr + rows, &pl->nnz, 8, cudaMemcpyHostToDevice, st));
    PG_CUDA_OK(cudaStreamSynchronize(st));  // &pl->nnz is pageable host memory
    return PG_OK;
}

int pg_plan_dof_permutation(const pg_plan *pl, int32_t *perm, void *stream) {
    PG_REQUIRE(pl && perm, PG_EINVAL, "pg_plan_dof_permutation: null pointer");
    dof_perm_kernel<<<(unsigned)((pl->nEnt + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        pl->nEnt, pl->p, pl->nE, pl->nF, pl->T, pl->blk_of_ent, pl->row_base, perm);
    PG_LAUNCH_OK();
    return PG_OK;
}

int64_t pg_plan_num_column_entities(const pg_plan *pl) { return pl ? pl->ncolent : -1; }

int pg_plan_column_starts(const pg_plan *pl, int32_t *out, void *stream) {
    PG_REQUIRE(pl && out, PG_EINVAL, "pg_plan_column_starts: null pointer");
    PG_CUDA_OK(cudaMemcpyAsync(out, pl->colstart, pl->ncolent * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return PG_OK;
}

int64_t pg_plan_entity_aligned_row(const pg_plan *pl, int64_t row) {
    if (!pl || row < 0) return -1;
    if (row >= pl->N) return pl->N;
    DevBuf res;
    int64_t hres[4];
    if (cudaMalloc(&res.p, 32) != cudaSuccess) return -1;
    find_blocks_kernel<<<1, 1>>>(pl->nEnt, pl->row_base, row, row, res.as<int64_t>());
    if (cudaMemcpy(hres, res.p, 32, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return hres[2];
}

}  // extern "C"

extern "C" int pg_plan_set_dirichlet(pg_plan *pl, const int32_t *elemsE, const int32_t *elemsF,
                                     const uint8_t *bd_entity, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_REQUIRE(pl, PG_EINVAL, "pg_plan_set_dirichlet: null plan");
    if (!bd_entity) {
        cudaFree(pl->bd_entity);
        pl->bd_entity = nullptr;
        return build_headers(pl, st);
    }
    PG_REQUIRE(elemsE && elemsF, PG_EINVAL, "pg_plan_set_dirichlet: null connectivity");
    if (!pl->bd_entity) PG_CUDA_OK(cudaMalloc((void **)&pl->bd_entity, pl->nEnt));
    PG_CUDA_OK(cudaMemcpyAsync(pl->bd_entity, bd_entity, pl->nEnt, cudaMemcpyDeviceToDevice, st));
    set_bdmask_kernel<<<(unsigned)((pl->nInc + 255) / 256), 256, 0, st>>>(pl->nInc, pl->nslots, elemsE, elemsF,
                                                                         pl->nE, pl->nF, pl->bd_entity, pl->rec);
    PG_LAUNCH_OK();
    return build_headers(pl, st);
}
