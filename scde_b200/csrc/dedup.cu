// dedup.cu -- the unique-count table indices (what scde.posteriors builds with unique()/match(),
// R/functions.R:631-632), on the device.
//
// Row ids are ascending in the count value within a cell -- a different order from R's first-appearance `unique()`,
// which is unobservable: rows are only ever addressed through the index matrix built here.
//
// Bitmap kernels (the common case: every count of the cell is below 65536).  One CTA per eight neighbouring cells.  Pass 1
// marks the cells' counts in eight 8 KB shared-memory bitmaps (a plain read before the atomicOr: after the first few
// genes almost every bit is already set), counts the bits and saves the bitmaps (80 MB for 10 000 cells).  Pass 2 reloads
// them, turns them into ranks (prefix population counts per word), writes the distinct values in ascending order and maps
// every gene's count to its row with two shared-memory reads and a popc; the eight cells' row ids of a gene are stored
// as one 32-byte sector of the gene-major index matrix ridx[g][c] (a CTA per cell would write 4 bytes per sector).
// HBM traffic: the counts are read once per pass, the index matrix is written once.
//
// Hash kernels (cells with a count >= 65536, flagged by pass 1): one CTA per cell; the column's distinct values are
// collected in a shared-memory open-addressing hash set, the set is sorted in place (bitonic, unsigned order so empty
// slots 0xFFFFFFFF sink to the end) and every gene's count is mapped to its rank by binary search.
#include "common.cuh"

namespace scde {
namespace {

constexpr int DEDUP_THREADS = 1024;
constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr int BM_WORDS = 2048;            // bitmap of the counts 0 .. 65535
constexpr int BM_STRIDE = BM_WORDS + 8;   // words per cell in the global bitmap scratch: the bitmap, then word BM_WORDS =
                                          // "hash kernel needed", word BM_WORDS + 1 = number of distinct counts >= 65536,
                                          // words BM_WORDS + 2 .. + 7 = those counts, ascending (at most OV_MAX)
constexpr int OV_MAX = 6;                 // distinct counts >= 65536 a cell may have and stay with the bitmap kernels
constexpr int OV_RAW = 32;                // occurrences collected in pass 1 before they are sorted and made unique
constexpr int BM_CPB = 8;                 // cells per CTA = one 32-byte sector of ridx per gene
constexpr int BM_THREADS = 512;

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
                   int32_t *__restrict__ n_unique, int32_t *err_flag, const uint32_t *__restrict__ bits) {
    extern __shared__ uint32_t s_tab[];
    __shared__ int s_count;
    const int c = blockIdx.x;
    if (bits && !bits[(size_t)c * BM_STRIDE + BM_WORDS]) return;  // the bitmap kernel has done this cell
    const int32_t *col = counts + (size_t)c * ldc + g0;
    build_set(col, G, s_tab, log2cap, err_flag, &s_count);
    if (threadIdx.x == 0) n_unique[c] = s_count;
}

__global__ void __launch_bounds__(DEDUP_THREADS)
dedup_emit_kernel(const int32_t *__restrict__ counts, int64_t ldc, int g0, int G, int n_cells, int log2cap_max,
                  const int32_t *__restrict__ row_off, int32_t *__restrict__ row_x, int32_t *__restrict__ ridx,
                  int ld_ridx, int32_t *err_flag, int64_t row_cap, const uint32_t *__restrict__ bits) {
    extern __shared__ uint32_t s_tab[];
    __shared__ int s_count;
    const int c = blockIdx.x;
    if (bits && !bits[(size_t)c * BM_STRIDE + BM_WORDS]) return;  // the bitmap kernel has done this cell
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

// ---- bitmap kernels --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BM_THREADS)
dedup_bitmap_count_kernel(const int32_t *__restrict__ counts, int64_t ldc, int g0, int G, int n_cells,
                          uint32_t *__restrict__ bits, int32_t *__restrict__ n_unique, int32_t *err_flag) {
    extern __shared__ uint32_t s_bm[];  // [BM_CPB][BM_WORDS]
    __shared__ int s_big[BM_CPB], s_ovn[BM_CPB];
    __shared__ uint32_t s_ov[BM_CPB][OV_RAW];
    const int c0 = blockIdx.x * BM_CPB;
    const int nc = min(BM_CPB, n_cells - c0);
    for (int i = threadIdx.x; i < BM_CPB * BM_WORDS; i += BM_THREADS) s_bm[i] = (i % BM_WORDS) == 0 ? 1u : 0u;  // count 0
    if (threadIdx.x < BM_CPB) s_big[threadIdx.x] = s_ovn[threadIdx.x] = 0;
    // EMPTY is not a count: a slot that is claimed but not yet written never matches
    for (int i = threadIdx.x; i < BM_CPB * OV_RAW; i += BM_THREADS) s_ov[i / OV_RAW][i % OV_RAW] = EMPTY;
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += BM_THREADS) {
        int32_t x[BM_CPB];
#pragma unroll
        for (int j = 0; j < BM_CPB; ++j) x[j] = j < nc ? counts[(size_t)(c0 + j) * ldc + g0 + g] : 0;
#pragma unroll
        for (int j = 0; j < BM_CPB; ++j) {
            if (x[j] < 0) {
                atomicOr(err_flag, 1);
            } else if (x[j] >= 32 * BM_WORDS) {
                // a handful of counts per cell at most: collected (with repeats) in a short list, sorted below
                const int n_seen = min(s_ovn[j], OV_RAW);
                bool seen = false;
                for (int i = 0; i < n_seen; ++i) seen = seen || s_ov[j][i] == (uint32_t)x[j];
                if (!seen) {
                    const int slot = atomicAdd(&s_ovn[j], 1);
                    if (slot < OV_RAW) s_ov[j][slot] = (uint32_t)x[j];
                    else s_big[j] = 1;
                }
            } else {
                uint32_t *w = s_bm + j * BM_WORDS + (x[j] >> 5);
                const uint32_t b = 1u << (x[j] & 31);
                if (!(*w & b)) atomicOr(w, b);
            }
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < nc; j += BM_THREADS / 32) {  // one warp per cell: population count, bitmap to global memory
        int n = 0;
        uint32_t *dst = bits + (size_t)(c0 + j) * BM_STRIDE;
        for (int w = lane; w < BM_WORDS; w += 32) {
            const uint32_t v = s_bm[j * BM_WORDS + w];
            n += __popc(v);
            dst[w] = v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        if (lane == 0) {
            // (a slot claimed by atomicAdd is written before the barrier above, so the list is complete here)
            int n_ov = 0;
            uint32_t ov[OV_RAW];
            if (!s_big[j]) {
                const int raw = min(s_ovn[j], OV_RAW);
                for (int i = 0; i < raw; ++i) {  // insertion sort, repeats dropped
                    const uint32_t v = s_ov[j][i];
                    int p = n_ov;
                    bool dup = false;
                    for (int q = 0; q < n_ov; ++q) dup = dup || ov[q] == v;
                    if (dup) continue;
                    while (p > 0 && ov[p - 1] > v) {
                        ov[p] = ov[p - 1];
                        --p;
                    }
                    ov[p] = v;
                    ++n_ov;
                }
                if (n_ov > OV_MAX) s_big[j] = 1;
            }
            dst[BM_WORDS] = (uint32_t)s_big[j];
            dst[BM_WORDS + 1] = s_big[j] ? 0u : (uint32_t)n_ov;
            for (int i = 0; i < OV_MAX; ++i) dst[BM_WORDS + 2 + i] = (!s_big[j] && i < n_ov) ? ov[i] : 0u;
            n_unique[c0 + j] = s_big[j] ? 0 : n + n_ov;  // flagged cells are counted by the hash kernel
        }
    }
}

__global__ void __launch_bounds__(BM_THREADS, 2)  // 96 KB of shared memory: two CTAs per SM, so at most 64 registers
dedup_bitmap_emit_kernel(const int32_t *__restrict__ counts, int64_t ldc, int g0, int G, int n_cells,
                         const uint32_t *__restrict__ bits, const int32_t *__restrict__ row_off,
                         int32_t *__restrict__ row_x, int32_t *__restrict__ ridx, int ld_ridx, int64_t row_cap) {
    extern __shared__ uint32_t s_bm[];                                               // [BM_CPB][BM_WORDS]
    uint16_t *s_pre = reinterpret_cast<uint16_t *>(s_bm + BM_CPB * BM_WORDS);        // [BM_CPB][BM_WORDS] ranks of the words
    __shared__ int s_part[BM_CPB][BM_THREADS / BM_CPB];
    __shared__ int s_base[BM_CPB], s_big[BM_CPB], s_ovn[BM_CPB], s_tot[BM_CPB];
    __shared__ uint32_t s_ov[BM_CPB][OV_MAX];
    const int c0 = blockIdx.x * BM_CPB;
    const int nc = min(BM_CPB, n_cells - c0);
    for (int i = threadIdx.x; i < BM_CPB * BM_WORDS; i += BM_THREADS) {
        const int j = i / BM_WORDS;
        s_bm[i] = j < nc ? bits[(size_t)(c0 + j) * BM_STRIDE + (i % BM_WORDS)] : 0u;
    }
    if (threadIdx.x < BM_CPB) {
        const int j = threadIdx.x;
        s_big[j] = j < nc ? (int)bits[(size_t)(c0 + j) * BM_STRIDE + BM_WORDS] : 1;
        s_ovn[j] = j < nc ? (int)bits[(size_t)(c0 + j) * BM_STRIDE + BM_WORDS + 1] : 0;
        for (int i = 0; i < OV_MAX; ++i) s_ov[j][i] = j < nc ? bits[(size_t)(c0 + j) * BM_STRIDE + BM_WORDS + 2 + i] : 0u;
        s_base[j] = j < nc ? row_off[c0 + j] : 0;
    }
    __syncthreads();
    // ranks: thread t owns WPT consecutive words of cell t / TPC
    constexpr int TPC = BM_THREADS / BM_CPB, WPT = BM_WORDS / TPC;
    const int cj = threadIdx.x / TPC, ct = threadIdx.x % TPC;
    const uint32_t *mine = s_bm + cj * BM_WORDS + ct * WPT;
    int local = 0;
#pragma unroll 4
    for (int w = 0; w < WPT; ++w) local += __popc(mine[w]);
    s_part[cj][ct] = local;
    __syncthreads();
    if (ct == 0) {  // exclusive scan over the cell's TPC partial sums
        int run = 0;
        for (int t = 0; t < TPC; ++t) {
            const int v = s_part[cj][t];
            s_part[cj][t] = run;
            run += v;
        }
        s_tot[cj] = run;
        // the counts >= 65536 follow the bitmap's values (they are larger than all of them), ascending
        if (cj < nc && !s_big[cj] && blockIdx.y == 0)
            for (int i = 0; i < s_ovn[cj]; ++i)
                if ((int64_t)s_base[cj] + run + i < row_cap) row_x[(int64_t)s_base[cj] + run + i] = (int32_t)s_ov[cj][i];
    }
    __syncthreads();
    {
        int run = s_part[cj][ct];
        const bool emit = cj < nc && !s_big[cj] && blockIdx.y == 0;
        const int64_t base = s_base[cj];
        for (int w = 0; w < WPT; ++w) {
            uint32_t v = mine[w];
            s_pre[cj * BM_WORDS + ct * WPT + w] = (uint16_t)run;
            while (v) {  // the distinct values in ascending order
                const int b = __ffs(v) - 1;
                v &= v - 1;
                if (emit && base + run < row_cap) row_x[base + run] = 32 * (ct * WPT + w) + b;
                ++run;
            }
        }
    }
    __syncthreads();
    const bool all8 = nc == BM_CPB && (ld_ridx % 4) == 0 && (reinterpret_cast<uintptr_t>(ridx + c0) % 16) == 0 && !s_big[0] &&
                      !s_big[1] && !s_big[2] && !s_big[3] && !s_big[4] && !s_big[5] && !s_big[6] && !s_big[7];
    // blockIdx.y: a slice of the genes (a launch over few cells -- one chunk of the pipelined front -- has too few groups of
    // eight cells to fill the GPU; the bitmaps and ranks are then rebuilt by every slice, which is cheap next to the mapping)
    const int g_per = (G + gridDim.y - 1) / gridDim.y;
    const int g_lo = blockIdx.y * g_per, g_hi = min(G, g_lo + g_per);
    for (int g = g_lo + threadIdx.x; g < g_hi; g += BM_THREADS) {
        int32_t x[BM_CPB], r[BM_CPB];
#pragma unroll
        for (int j = 0; j < BM_CPB; ++j) x[j] = j < nc ? counts[(size_t)(c0 + j) * ldc + g0 + g] : 0;
        bool big = false;
#pragma unroll
        for (int j = 0; j < BM_CPB; ++j) {
            const uint32_t xv = x[j] < 0 ? 0u : (uint32_t)x[j];
            const uint32_t w = min(xv >> 5, (uint32_t)(BM_WORDS - 1));
            r[j] = s_base[j] + (int)s_pre[j * BM_WORDS + w] + __popc(s_bm[j * BM_WORDS + w] & ((1u << (xv & 31)) - 1u));
            big = big || xv >= 32u * BM_WORDS;
        }
        if (big) {  // rare: some cell's count lies beyond the bitmap -- it is one of the cell's few listed values
#pragma unroll
            for (int j = 0; j < BM_CPB; ++j) {
                const int32_t xr = j < nc ? counts[(size_t)(c0 + j) * ldc + g0 + g] : 0;  // re-read: keeps the hot loop lean
                const uint32_t xv = xr < 0 ? 0u : (uint32_t)xr;
                if (xv >= 32u * BM_WORDS) {
                    int i = 0;
                    while (i < OV_MAX - 1 && s_ov[j][i] != xv) ++i;
                    r[j] = s_base[j] + s_tot[j] + i;
                }
            }
        }
        int32_t *dst = ridx + (size_t)g * ld_ridx + c0;
        if (all8) {
            reinterpret_cast<int4 *>(dst)[0] = make_int4(r[0], r[1], r[2], r[3]);
            reinterpret_cast<int4 *>(dst)[1] = make_int4(r[4], r[5], r[6], r[7]);
        } else {
#pragma unroll
            for (int j = 0; j < BM_CPB; ++j)
                if (j < nc && !s_big[j]) dst[j] = r[j];
        }
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

size_t dedup_scratch_words(int n_cells) { return (size_t)(n_cells > 0 ? n_cells : 1) * BM_STRIDE; }

cudaError_t launch_dedup_count(const int32_t *counts, int64_t ld_counts, int g0, int G, int n_cells,
                               int32_t *n_unique, int32_t *err_flag, uint32_t *scratch, cudaStream_t st) {
    if (n_cells <= 0) return cudaSuccess;
    const size_t bsmem = sizeof(uint32_t) * BM_CPB * BM_WORDS;
    cudaError_t e = cudaFuncSetAttribute(dedup_bitmap_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem);
    if (e != cudaSuccess) return e;
    dedup_bitmap_count_kernel<<<(n_cells + BM_CPB - 1) / BM_CPB, BM_THREADS, bsmem, st>>>(counts, ld_counts, g0, G, n_cells,
                                                                                          scratch, n_unique, err_flag);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // cells with a count >= 65536 (flagged in the scratch): hash kernel, every other CTA leaves at once
    int l = pick_log2cap(G);
    size_t smem = sizeof(uint32_t) << l;
    e = cudaFuncSetAttribute(dedup_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dedup_count_kernel<<<n_cells, DEDUP_THREADS, smem, st>>>(counts, ld_counts, g0, G, n_cells, l, n_unique, err_flag, scratch);
    return cudaGetLastError();
}

cudaError_t launch_exclusive_scan(const int32_t *in, int32_t *out, int n, const int32_t *base, cudaStream_t st) {
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(in, out, n, base);
    return cudaGetLastError();
}

cudaError_t launch_dedup_emit(const int32_t *counts, int64_t ld_counts, int g0, int G, int n_cells,
                              const int32_t *row_off, int32_t *row_x, int32_t *ridx, int ld_ridx,
                              int32_t *err_flag, int64_t row_cap, const uint32_t *scratch, cudaStream_t st) {
    if (n_cells <= 0) return cudaSuccess;
    const size_t bsmem = (sizeof(uint32_t) + sizeof(uint16_t)) * BM_CPB * BM_WORDS;
    cudaError_t e = cudaFuncSetAttribute(dedup_bitmap_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem);
    if (e != cudaSuccess) return e;
    const int groups = (n_cells + BM_CPB - 1) / BM_CPB;
    int slices = (2 * 148 + groups - 1) / groups;
    slices = slices < 1 ? 1 : (slices > 4 ? 4 : slices);
    dedup_bitmap_emit_kernel<<<dim3(groups, slices), BM_THREADS, bsmem, st>>>(
        counts, ld_counts, g0, G, n_cells, scratch, row_off, row_x, ridx, ld_ridx, row_cap);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int l = pick_log2cap(G);
    size_t smem = sizeof(uint32_t) << l;
    e = cudaFuncSetAttribute(dedup_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dedup_emit_kernel<<<n_cells, DEDUP_THREADS, smem, st>>>(counts, ld_counts, g0, G, n_cells, l, row_off, row_x, ridx,
                                                            ld_ridx, err_flag, row_cap, scratch);
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
