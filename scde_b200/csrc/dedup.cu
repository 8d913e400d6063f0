// dedup.cu -- the unique-count table indices (what scde.posteriors builds with unique()/match(),
// R/functions.R:631-632), on the device.
//
// One CTA per cell (one column of the count matrix).  The column's distinct values are collected in a shared
// memory open-addressing hash set, the set is sorted in place (bitonic, unsigned order so empty slots
// 0xFFFFFFFF sink to the end) and every gene's count is mapped to its rank by binary search.  Row ids are
// ascending in the count value within a cell -- a different order from R's first-appearance `unique()`, which
// is unobservable: rows are only ever addressed through the index matrix built here.
//
// HBM traffic: the count column is read twice per pass (coalesced), the index matrix is written once
// gene-major (ridx[g][c]) so the contraction kernel reads one contiguous run per gene.
#include "common.cuh"

namespace scde {
namespace {

constexpr int DEDUP_THREADS = 1024;
constexpr uint32_t EMPTY = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t hash_slot(uint32_t x, int log2cap) { return (x * 2654435761u) >> (32 - log2cap); }

// inserts all counts of the column into s_tab (capacity cap = 1 << log2cap); returns false on capacity overflow
__device__ bool build_set(const int32_t *__restrict__ col, int G, uint32_t *s_tab, int log2cap, int32_t *err_flag,
                          int *s_count) {
    const uint32_t cap = 1u << log2cap, mask = cap - 1;
    for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) s_tab[i] = EMPTY;
    if (threadIdx.x == 0) *s_count = 0;
    __syncthreads();
    // g == -1 stands for the count 0, which every cell gets a table row for whether or not a gene has it (the zero-count
    // row is the base of the zero-base contraction)
    for (int g = (int)threadIdx.x - 1; g < G; g += blockDim.x) {
        int32_t xi = g < 0 ? 0 : col[g];
        if (xi < 0) {
            atomicOr(err_flag, 1);
            continue;
        }
        uint32_t x = (uint32_t)xi;
        uint32_t h = hash_slot(x, log2cap);
        for (uint32_t probe = 0; probe < cap; ++probe) {
            uint32_t cur = s_tab[h];
            if (cur == x) break;
            if (cur == EMPTY) {
                uint32_t prev = atomicCAS(&s_tab[h], EMPTY, x);
                if (prev == EMPTY) {
                    atomicAdd(s_count, 1);
                    break;
                }
                if (prev == x) break;
            }
            h = (h + 1) & mask;
        }
    }
    __syncthreads();
    // one slot must stay empty so that unsuccessful probes terminate
    if (*s_count >= (int)cap) {
        if (threadIdx.x == 0) atomicOr(err_flag, 2);
        return false;
    }
    return true;
}

__device__ void bitonic_sort(uint32_t *s, uint32_t n) {  // n a power of two; ascending unsigned
    for (uint32_t k = 2; k <= n; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                uint32_t l = i ^ j;
                if (l > i) {
                    uint32_t a = s[i], b = s[l];
                    bool up = ((i & k) == 0);
                    if ((a > b) == up) {
                        s[i] = b;
                        s[l] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(DEDUP_THREADS)
dedup_count_kernel(const int32_t *__restrict__ counts, int64_t ldc, int g0, int G, int n_cells, int log2cap,
                   int32_t *__restrict__ n_unique, int32_t *err_flag) {
    extern __shared__ uint32_t s_tab[];
    __shared__ int s_count;
    const int c = blockIdx.x;
    const int32_t *col = counts + (size_t)c * ldc + g0;
    build_set(col, G, s_tab, log2cap, err_flag, &s_count);
    if (threadIdx.x == 0) n_unique[c] = s_count;
}

__global__ void __launch_bounds__(DEDUP_THREADS)
dedup_emit_kernel(const int32_t *__restrict__ counts, int64_t ldc, int g0, int G, int n_cells, int log2cap_max,
                  const int32_t *__restrict__ row_off, int32_t *__restrict__ row_x, int32_t *__restrict__ ridx,
                  int ld_ridx, int32_t *err_flag, int64_t row_cap) {
    extern __shared__ uint32_t s_tab[];
    __shared__ int s_count;
    const int c = blockIdx.x;
    const int32_t *col = counts + (size_t)c * ldc + g0;
    // the count pass told us how many distinct values this cell has: size the set (and the sort) for that, not for G
    const int n_distinct = row_off[c + 1] - row_off[c];
    int log2cap = 5;
    while ((1 << log2cap) < 2 * n_distinct + 2 && log2cap < log2cap_max) ++log2cap;
    if (!build_set(col, G, s_tab, log2cap, err_flag, &s_count)) return;
    const int U = s_count;
    bitonic_sort(s_tab, 1u << log2cap);
    const int base = row_off[c];
    for (int i = threadIdx.x; i < U; i += blockDim.x)
        if ((int64_t)base + i < row_cap) row_x[base + i] = (int32_t)s_tab[i];  // beyond the capacity: api.cu falls back
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        int32_t xi = col[g];
        uint32_t x = xi < 0 ? 0u : (uint32_t)xi;
        int lo = 0, hi = U - 1;  // x is present
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (s_tab[mid] < x) lo = mid + 1; else hi = mid;
        }
        ridx[(size_t)g * ld_ridx + c] = base + lo;
    }
}

__global__ void exclusive_scan_kernel(const int32_t *__restrict__ in, int32_t *__restrict__ out, int n,
                                      const int32_t *__restrict__ base_ptr) {
    // single CTA, 1024 threads, chunked Hillis-Steele over n elements; out[n] = base + total
    __shared__ int32_t s[1024];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = base_ptr ? *base_ptr : 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        int i = base + threadIdx.x;
        int32_t v = i < n ? in[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int32_t t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n) out[i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

__global__ void uci_to_ridx_kernel(const int32_t *__restrict__ uci, int G, int n_cells,
                                   const int32_t *__restrict__ ucl_off, int32_t *__restrict__ ridx, int ld_ridx) {
    // tiled transpose: uci is [c][g] in memory (column-major G x C), ridx is [g][c]
    __shared__ int32_t tile[32][33];
    int gx = blockIdx.x * 32, cy = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int c = cy + j, g = gx + threadIdx.x;
        if (c < n_cells && g < G) tile[j][threadIdx.x] = ucl_off[c] + uci[(size_t)c * G + g];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int g = gx + j, c = cy + threadIdx.x;
        if (c < n_cells && g < G) ridx[(size_t)g * ld_ridx + c] = tile[threadIdx.x][j];
    }
}

int pick_log2cap(int G) {
    // capacity > distinct values is required; distinct <= G + 1 (the forced 0).  Cap at 2^15 slots (128 KB).
    int l = 5;
    while ((1 << l) <= G + 1 && l < 15) ++l;
    if ((1 << l) < 2 * G && l < 15) ++l;  // keep the load factor below 1/2 when that is free
    return l;
}

}  // namespace

cudaError_t launch_dedup_count(const int32_t *counts, int64_t ld_counts, int g0, int G, int n_cells,
                               int32_t *n_unique, int32_t *err_flag, cudaStream_t st) {
    if (n_cells <= 0) return cudaSuccess;
    int l = pick_log2cap(G);
    size_t smem = sizeof(uint32_t) << l;
    cudaError_t e = cudaFuncSetAttribute(dedup_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dedup_count_kernel<<<n_cells, DEDUP_THREADS, smem, st>>>(counts, ld_counts, g0, G, n_cells, l, n_unique, err_flag);
    return cudaGetLastError();
}

cudaError_t launch_exclusive_scan(const int32_t *in, int32_t *out, int n, const int32_t *base, cudaStream_t st) {
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(in, out, n, base);
    return cudaGetLastError();
}

cudaError_t launch_dedup_emit(const int32_t *counts, int64_t ld_counts, int g0, int G, int n_cells,
                              const int32_t *row_off, int32_t *row_x, int32_t *ridx, int ld_ridx,
                              int32_t *err_flag, int64_t row_cap, cudaStream_t st) {
    if (n_cells <= 0) return cudaSuccess;
    int l = pick_log2cap(G);
    size_t smem = sizeof(uint32_t) << l;
    cudaError_t e = cudaFuncSetAttribute(dedup_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dedup_emit_kernel<<<n_cells, DEDUP_THREADS, smem, st>>>(counts, ld_counts, g0, G, n_cells, l, row_off, row_x, ridx,
                                                            ld_ridx, err_flag, row_cap);
    return cudaGetLastError();
}

cudaError_t launch_uci_to_ridx(const int32_t *uci, int G, int n_cells, const int32_t *ucl_off, int32_t *ridx,
                               int ld_ridx, cudaStream_t st) {
    if (G <= 0 || n_cells <= 0) return cudaSuccess;
    dim3 grid((G + 31) / 32, (n_cells + 31) / 32), block(32, 8);
    uci_to_ridx_kernel<<<grid, block, 0, st>>>(uci, G, n_cells, ucl_off, ridx, ld_ridx);
    return cudaGetLastError();
}

}  // namespace scde
