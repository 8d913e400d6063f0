// boot_contract.cu -- the bootstrap joint posterior (src/jpmatLogBoot.cpp:216-275, :460-499, :11-86).
//
// The reference loops  for b < B: for each draw r: for g < G: tjp[:, g] += lp_row(r, g)  and then soft-maxes
// every gene column and adds tjp/(sum*B) into jp.  Per gene this is a contraction
//     T[b, k] = sum_c W[c, b] * lp[c, x[g, c], k],       jp[g, k] = (1/B) sum_b softmax_k(T[b, :])
// with W the cell-by-randomization multiplicity matrix (W[c, b] = how many of boot b's draws hit cell c).
// K = 401 grid points, B = 100 randomizations, FP64 (lp spans [-751, 0] plus a -1.6e304 "log 0" sentinel and a
// 1e-6 relative tolerance on log-posteriors leaves no room for a reduced-precision tensor-core split in this round;
// tcgen05 has no FP64 kind).
//
// contract_tiled_kernel (the sm_100a hot kernel)
//   * one 2-CTA cluster per gene; CTA r owns grid points [208 r, 208 r + 208) and all 104 (100 + pad) boots, so
//     its 208 x 104 FP64 accumulator tile (173 KB) lives entirely in the register file (12 warps x 168 regs);
//   * operands are staged through shared memory by the TMA engine: per stage of 8 cells, 8 bulk copies of one
//     gathered 1664-byte table row half each plus one bulk copy of the 8 matching W rows, completion signalled on
//     an mbarrier (cp.async.bulk ... mbarrier::complete_tx); an 8-deep ring keeps ~150 KB in flight per SM;
//   * warps are laid out so that every SM sub-partition holds the same number of accumulators
//     (three warps with 4x13, 4x13 and 5x13 register tiles = 52 grid points x 104 boots each) -- the FP64 pipe is
//     per sub-partition, so an unbalanced split would idle a quarter of it;
//   * lanes are 4 (grid) x 8 (boots): a warp reads 16..20 consecutive doubles of the table row and all 104 W
//     values per cell, i.e. ~9 shared-memory wavefronts for 52..65 DFMA warp-instructions;
//   * the soft-max over the grid is fused: per-boot max and sum are reduced with warp shuffles, across warps
//     through shared memory and across the two CTAs through distributed shared memory, then every thread adds
//     exp(T - max)/(sum * B) over its boots and writes jp -- T never leaves the chip.
// contract_generic_kernel handles any K / any B (used for K > 416 and as an on-device cross-check).
#include "common.cuh"
#include <cfloat>
#include <cmath>

namespace scde {
namespace {

// ------------------------------------------------------------------------------------------------
// small PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// TMA bulk copy global -> shared (this CTA), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t n_clusters_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
// store a double into the peer CTA's shared memory at the same offset as local pointer `p`
__device__ __forceinline__ void st_peer_f64(const void *p, uint32_t peer, double v) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(peer));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// W build: multiplicities from the draw lists
__global__ void build_w_kernel(const int32_t *__restrict__ boot_idx, int n_boot, int D, int n_list, double *W,
                               int n_w_rows) {
    // W is pass-major: W[pass][cell][104], boot b = 104*pass + column.  One CTA per boot, so atomics from
    // different CTAs never touch the same element.
    const int b = blockIdx.x;
    double *Wp = W + ((size_t)(b / WP_TILED) * n_w_rows) * WP_TILED + (b % WP_TILED);
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        int c = boot_idx[(size_t)b * D + j];
        if (c >= 0 && c < n_list) atomicAdd(&Wp[(size_t)c * WP_TILED], 1.0);
    }
}

// ------------------------------------------------------------------------------------------------
// tiled kernel
constexpr int T_KH = 208;       // grid points per CTA
constexpr int T_WP = WP_TILED;  // boots per pass (104)
constexpr int T_S = 8;          // cells per stage
constexpr int T_NS = 8;         // ring depth
constexpr int T_PD = T_NS - 2;  // prefetch distance: the slot refilled at iteration i was consumed at i-2
constexpr int T_WARPS = 12;
constexpr int T_THREADS = T_WARPS * 32;
constexpr int T_STAGE_A = T_S * T_KH;                      // doubles
constexpr int T_STAGE_W = T_S * T_WP;                      // doubles
constexpr int T_STAGE_DOUBLES = T_STAGE_A + T_STAGE_W;     // 2496
constexpr uint32_t T_STAGE_BYTES = T_STAGE_DOUBLES * 8u;   // 19968
constexpr int T_NB = 13;                                   // boots per thread

struct TiledSmem {
    double stage[T_NS][T_STAGE_DOUBLES];
    double red[T_WARPS][T_WP];
    double xmax[2][T_WP];  // [0] this CTA's value, [1] written by the peer
    double xsum[2][T_WP];
    uint64_t full[T_NS];
    uint64_t empty[T_NS];
};

struct TiledParams {
    const double *table;
    const int32_t *ridx;
    int64_t ld_ridx;
    const int32_t *cell_ids;
    int n_list;
    const double *W;  // this pass: rows [n_w_rows][ldw], columns [0, 104)
    int64_t ldw;
    int n_boot_pass;  // real boots in this pass (<= 104)
    double scale;
    int n_genes, K;
    double *jp;
    int64_t ld_jp;
    int accumulate;
};

template <int TK>
__device__ __forceinline__ void consume_stage(const double *__restrict__ sA, const double *__restrict__ sW,
                                              double (&acc)[TK][T_NB], int a_off, int a4_off, int lb) {
#pragma unroll 2
    for (int c = 0; c < T_S; ++c) {
        const double *a = sA + c * T_KH + a_off;
        const double2 a01 = *reinterpret_cast<const double2 *>(a);
        const double2 a23 = *reinterpret_cast<const double2 *>(a + 2);
        double av[TK];
        av[0] = a01.x;
        av[1] = a01.y;
        av[2] = a23.x;
        av[3] = a23.y;
        if (TK == 5) av[TK - 1] = sA[c * T_KH + a4_off];
        const double *w = sW + c * T_WP + 2 * lb;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const double2 wv = *reinterpret_cast<const double2 *>(w + 16 * j);
#pragma unroll
            for (int i = 0; i < TK; ++i) {
                acc[i][2 * j] = fma(av[i], wv.x, acc[i][2 * j]);
                acc[i][2 * j + 1] = fma(av[i], wv.y, acc[i][2 * j + 1]);
            }
        }
        const double w12 = sW[c * T_WP + 96 + lb];
#pragma unroll
        for (int i = 0; i < TK; ++i) acc[i][12] = fma(av[i], w12, acc[i][12]);
    }
}

// boot column of accumulator slot j for boot-lane lb
__device__ __forceinline__ int boot_of(int j, int lb) { return j < 12 ? 16 * (j >> 1) + 2 * lb + (j & 1) : 96 + lb; }

template <int TK>
__device__ __forceinline__ void run_tiles(const TiledParams &p, TiledSmem &sm, int warp, int lane, uint32_t rank,
                                          int n_my_genes, int spg, int kw) {
    const int lk = lane & 3, lb = lane >> 2;
    const int a_off = kw + lk * 4;    // first of the thread's 4 consecutive grid points (CTA-relative)
    const int a4_off = kw + 16 + lk;  // fifth grid point of the 5-wide tiles
    const int kbase = rank * T_KH;
    const uint32_t peer = rank ^ 1u;
    const int64_t total_stages = (int64_t)n_my_genes * spg;
    const uint32_t cid = cluster_id_x(), ncl = n_clusters_x();

    // ---- producer state (warp 0 only): lanes 0..7 gather table rows, lane 8 copies the W rows ----
    int64_t pq = 0;             // next stage to issue
    int p_gi = 0, p_cb = 0;     // its (gene ordinal, cell block)
    int32_t next_row = 0;       // row index prefetched for stage pq (lane j: cell p_cb*8 + j)
    auto prefetch_row = [&](int gi, int cb) -> int32_t {
        if (lane < T_S && gi < n_my_genes) {
            int cell = cb * T_S + lane;
            if (cell >= p.n_list) cell = p.n_list - 1;  // padded cells re-read a valid row; their W rows are zero
            int col = p.cell_ids ? p.cell_ids[cell] : cell;
            int64_t gene = (int64_t)cid + (int64_t)gi * ncl;
            return p.ridx[gene * p.ld_ridx + col];
        }
        return 0;
    };
    auto issue_stage = [&]() {  // issues stage pq using next_row, then prefetches the following stage's rows
        const int slot = (int)(pq % T_NS);
        const uint32_t fill = (uint32_t)(pq / T_NS);
        if (fill > 0) mbar_wait(&sm.empty[slot], (fill - 1) & 1u);
        double *dstA = sm.stage[slot];
        double *dstW = dstA + T_STAGE_A;
        if (lane == 0) mbar_arrive_expect_tx(&sm.full[slot], T_STAGE_BYTES);
        __syncwarp();
        if (lane < T_S) {
            bulk_g2s(dstA + lane * T_KH, p.table + (int64_t)next_row * KP_TILED + kbase, T_KH * 8u, &sm.full[slot]);
        } else if (lane == T_S) {
            bulk_g2s(dstW, p.W + (int64_t)p_cb * T_S * p.ldw, T_STAGE_W * 8u, &sm.full[slot]);
        }
        ++pq;
        if (++p_cb == spg) {
            p_cb = 0;
            ++p_gi;
        }
        next_row = prefetch_row(p_gi, p_cb);
    };
    if constexpr (TK == 4) {  // warp 0 always runs the 4-wide instantiation
        if (warp == 0) {
            next_row = prefetch_row(0, 0);
            for (int i = 0; i < T_PD && pq < total_stages; ++i) issue_stage();
        }
    }

    int64_t q = 0;  // stage being consumed
    for (int gi = 0; gi < n_my_genes; ++gi) {
        const int64_t gene = (int64_t)cid + (int64_t)gi * ncl;
        double acc[TK][T_NB];
#pragma unroll
        for (int i = 0; i < TK; ++i)
#pragma unroll
            for (int j = 0; j < T_NB; ++j) acc[i][j] = 0.0;

        for (int cb = 0; cb < spg; ++cb, ++q) {
            if constexpr (TK == 4) {
                if (warp == 0 && pq < total_stages) issue_stage();
            }
            const int slot = (int)(q % T_NS);
            mbar_wait(&sm.full[slot], (uint32_t)(q / T_NS) & 1u);
            const double *sA = sm.stage[slot];
            consume_stage<TK>(sA, sA + T_STAGE_A, acc, a_off, a4_off, lb);
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[slot]);
        }

        // ---------------- fused soft-max over the grid and average over boots ----------------
        bool kvalid[TK];
#pragma unroll
        for (int i = 0; i < TK; ++i) {
            int kk = (i < 4) ? a_off + i : a4_off;
            kvalid[i] = (kbase + kk) < p.K;
        }
        // (1) per-boot maximum over this CTA's grid points
#pragma unroll
        for (int j = 0; j < T_NB; ++j) {
            double m = -INFINITY;
#pragma unroll
            for (int i = 0; i < TK; ++i)
                if (kvalid[i]) m = fmax(m, acc[i][j]);
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
            if (lk == 0) sm.red[warp][boot_of(j, lb)] = m;
        }
        named_bar_sync(1, T_THREADS);
        if (threadIdx.x < T_WP) {
            double m = sm.red[0][threadIdx.x];
#pragma unroll
            for (int w = 1; w < T_WARPS; ++w) m = fmax(m, sm.red[w][threadIdx.x]);
            sm.xmax[0][threadIdx.x] = m;
            st_peer_f64(&sm.xmax[1][threadIdx.x], peer, m);
        }
        cluster_arrive();
        cluster_wait();
        // (2) exponentials and per-boot sums
#pragma unroll
        for (int j = 0; j < T_NB; ++j) {
            const int b = boot_of(j, lb);
            const double M = fmax(sm.xmax[0][b], sm.xmax[1][b]);
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < TK; ++i) {
                double e = kvalid[i] ? exp(acc[i][j] - M) : 0.0;
                acc[i][j] = e;
                s += e;
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (lk == 0) sm.red[warp][b] = s;
        }
        named_bar_sync(1, T_THREADS);
        if (threadIdx.x < T_WP) {
            double s = sm.red[0][threadIdx.x];
#pragma unroll
            for (int w = 1; w < T_WARPS; ++w) s += sm.red[w][threadIdx.x];
            sm.xsum[0][threadIdx.x] = s;
            st_peer_f64(&sm.xsum[1][threadIdx.x], peer, s);
        }
        cluster_arrive();
        cluster_wait();
        // (3) jp[g, k] += sum_b e[k, b] / (S_b * scale)
        double r[TK];
#pragma unroll
        for (int i = 0; i < TK; ++i) r[i] = 0.0;
#pragma unroll
        for (int j = 0; j < T_NB; ++j) {
            const int b = boot_of(j, lb);
            if (b < p.n_boot_pass) {
                // rank-0 value first so both CTAs add in the same order
                const double s0 = rank == 0 ? sm.xsum[0][b] : sm.xsum[1][b];
                const double s1 = rank == 0 ? sm.xsum[1][b] : sm.xsum[0][b];
                const double den = (s0 + s1) * p.scale;
#pragma unroll
                for (int i = 0; i < TK; ++i) r[i] += acc[i][j] / den;
            }
        }
#pragma unroll
        for (int i = 0; i < TK; ++i) {
            r[i] += __shfl_xor_sync(0xffffffffu, r[i], 4);
            r[i] += __shfl_xor_sync(0xffffffffu, r[i], 8);
            r[i] += __shfl_xor_sync(0xffffffffu, r[i], 16);
        }
        if (lb == 0) {
            double *out = p.jp + gene * p.ld_jp + kbase;
#pragma unroll
            for (int i = 0; i < TK; ++i) {
                int kk = (i < 4) ? a_off + i : a4_off;
                if (kvalid[i]) {
                    if (p.accumulate) out[kk] += r[i]; else out[kk] = r[i];
                }
            }
        }
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T_THREADS, 1) contract_tiled_kernel(const TiledParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TiledSmem &sm = *reinterpret_cast<TiledSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t cid = cluster_id_x(), ncl = n_clusters_x();
    if (threadIdx.x == 0) {
        for (int s = 0; s < T_NS; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], T_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // both CTAs must be resident before any DSMEM store
    cluster_arrive();
    cluster_wait();

    const int n_my_genes = ((int)cid < p.n_genes) ? (p.n_genes - (int)cid + (int)ncl - 1) / (int)ncl : 0;
    const int spg = (p.n_list + T_S - 1) / T_S;
    // warp -> sub-partition (warp & 3) and slot (warp >> 2): slots 0,1 own 16 grid points, slot 2 owns 20
    const int smsp = warp & 3, slot = warp >> 2;
    const int kw = smsp * 52 + slot * 16;
    if (slot == 2)
        run_tiles<5>(p, sm, warp, lane, rank, n_my_genes, spg, kw);
    else
        run_tiles<4>(p, sm, warp, lane, rank, n_my_genes, spg, kw);
    // keep this CTA's shared memory alive until the peer's last DSMEM store has landed
    cluster_arrive();
    cluster_wait();
}

// ------------------------------------------------------------------------------------------------
// tiled kernel, DMMA form (the default).  Same cluster / TMA ring / soft-max structure as above, but the inner
// product runs on mma.sync.aligned.m8n8k4.f64: M = 8 grid points, N = 8 boots, K = 4 cells per instruction.  The FP64
// rate of DMMA equals that of DFMA on B200 (37 vs 36.5 TFLOP/s measured, tools/microbench.cu), but one DMMA replaces
// eight DFMA warp-instructions and its fragments are one double per lane, so a warp issues 15 LDS.64 + 26 DMMA per four
// cells instead of 36 LDS + 208 DFMA -- the register-tile version was limited by shared-memory instruction issue
// (LDS.128 sustains one per two cycles per SM) and by issue slots, not by the FP64 pipe (profiles/r01a_*).
//
// Tiles per CTA: 26 (grid) x 13 (boots).  Sub-partition s (= warp & 3) owns grid tiles 6s..6s+5 completely -- two per
// warp slot -- and half of a shared grid tile (24 for s = 0,1; 25 for s = 2,3): boot tiles 0..6 for even s, 6..12 for
// odd s, where the duplicated boot tile 6 of the odd warps is computed but masked out.  That is 85 tiles per
// sub-partition, i.e. the FP64 pipe of every sub-partition carries the same load.
// Shared-memory rows are padded to 216 doubles so the four cells of an A fragment fall into disjoint bank halves
// (216 * 2 words = 16 mod 32), as the 104-double W rows already do: every fragment load is conflict-free.
constexpr int M_AS = 216;                                   // padded row stride of the A stage (doubles)
constexpr int M_STAGE_DOUBLES = T_S * M_AS + T_S * T_WP;    // 2560
constexpr int M_NT = 13;                                    // boot tiles

struct MmaSmem {
    double stage[T_NS][M_STAGE_DOUBLES];
    double red[T_WARPS][T_WP];
    double xmax[2][T_WP];
    double xsum[2][T_WP];
    uint64_t full[T_NS];
    uint64_t empty[T_NS];
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// NEX: extra tiles of the shared grid tile (0 or 7); NX0: first boot tile of the extras (0 or 6)
template <int NEX, int NX0>
__device__ __forceinline__ void run_mma(const TiledParams &p, MmaSmem &sm, int warp, int lane, uint32_t rank,
                                        int n_my_genes, int spg) {
    const int g = lane >> 2, t = lane & 3;
    const int smsp = warp & 3, slot = warp >> 2;
    const int m0 = smsp * 6 + slot * 2;   // first of the two full grid tiles
    const int mx = 24 + (smsp >> 1);      // shared grid tile (slot 2 only)
    const int kbase = rank * T_KH;
    const uint32_t peer = rank ^ 1u;
    const int64_t total_stages = (int64_t)n_my_genes * spg;
    const uint32_t cid = cluster_id_x(), ncl = n_clusters_x();

    // ---- producer (warp 0; it always runs the NEX == 0 instantiation) ----
    int64_t pq = 0;
    int p_gi = 0, p_cb = 0;
    int32_t next_row = 0;
    auto prefetch_row = [&](int gi, int cb) -> int32_t {
        if (lane < T_S && gi < n_my_genes) {
            int cell = cb * T_S + lane;
            if (cell >= p.n_list) cell = p.n_list - 1;
            int col = p.cell_ids ? p.cell_ids[cell] : cell;
            int64_t gene = (int64_t)cid + (int64_t)gi * ncl;
            return p.ridx[gene * p.ld_ridx + col];
        }
        return 0;
    };
    auto issue_stage = [&]() {
        const int sl = (int)(pq % T_NS);
        const uint32_t fill = (uint32_t)(pq / T_NS);
        if (fill > 0) mbar_wait(&sm.empty[sl], (fill - 1) & 1u);
        double *dstA = sm.stage[sl];
        double *dstW = dstA + T_S * M_AS;
        if (lane == 0) mbar_arrive_expect_tx(&sm.full[sl], T_STAGE_BYTES);
        __syncwarp();
        if (lane < T_S) {
            bulk_g2s(dstA + lane * M_AS, p.table + (int64_t)next_row * KP_TILED + kbase, T_KH * 8u, &sm.full[sl]);
        } else if (lane == T_S) {
            bulk_g2s(dstW, p.W + (int64_t)p_cb * T_S * T_WP, T_STAGE_W * 8u, &sm.full[sl]);
        }
        ++pq;
        if (++p_cb == spg) {
            p_cb = 0;
            ++p_gi;
        }
        next_row = prefetch_row(p_gi, p_cb);
    };
    if constexpr (NEX == 0) {
        if (warp == 0) {
            next_row = prefetch_row(0, 0);
            for (int i = 0; i < T_PD && pq < total_stages; ++i) issue_stage();
        }
    }

    // validity of this thread's grid points
    const bool kv0 = (kbase + (m0 + 0) * 8 + g) < p.K;
    const bool kv1 = (kbase + (m0 + 1) * 8 + g) < p.K;
    const bool kvx = (kbase + mx * 8 + g) < p.K;

    int64_t q = 0;
    for (int gi = 0; gi < n_my_genes; ++gi) {
        const int64_t gene = (int64_t)cid + (int64_t)gi * ncl;
        double acc[2][M_NT][2];
        double ex[NEX > 0 ? NEX : 1][2];
#pragma unroll
        for (int nt = 0; nt < M_NT; ++nt) {
            acc[0][nt][0] = acc[0][nt][1] = 0.0;
            acc[1][nt][0] = acc[1][nt][1] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < (NEX > 0 ? NEX : 1); ++j) ex[j][0] = ex[j][1] = 0.0;

        for (int cb = 0; cb < spg; ++cb, ++q) {
            if constexpr (NEX == 0) {
                if (warp == 0 && pq < total_stages) issue_stage();
            }
            const int sl = (int)(q % T_NS);
            mbar_wait(&sm.full[sl], (uint32_t)(q / T_NS) & 1u);
            const double *sA = sm.stage[sl];
            const double *sW = sA + T_S * M_AS;
#pragma unroll
            for (int ks = 0; ks < T_S / 4; ++ks) {
                const double *ap = sA + (ks * 4 + t) * M_AS + g;
                const double *wp = sW + (ks * 4 + t) * T_WP + g;
                const double a0 = ap[(m0 + 0) * 8];
                const double a1 = ap[(m0 + 1) * 8];
                double a2 = 0.0;
                if constexpr (NEX > 0) a2 = ap[mx * 8];
#pragma unroll
                for (int nt = 0; nt < M_NT; ++nt) {
                    const double b = wp[nt * 8];
                    dmma(acc[0][nt], a0, b);
                    dmma(acc[1][nt], a1, b);
                    if constexpr (NEX > 0) {
                        if (nt >= NX0 && nt < NX0 + NEX) dmma(ex[nt - NX0], a2, b);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[sl]);
        }

        // ---------------- fused soft-max over the grid and average over boots ----------------
        // thread holds, per tile, C[m = g][n = 2t + i]: grid point (tile*8 + g), boot (nt*8 + 2t + i)
        // (1) per-boot maximum over this CTA's grid points
#pragma unroll
        for (int nt = 0; nt < M_NT; ++nt) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                double m = -INFINITY;
                if (kv0) m = fmax(m, acc[0][nt][i]);
                if (kv1) m = fmax(m, acc[1][nt][i]);
                if constexpr (NEX > 0) {
                    if (nt >= NX0 && nt < NX0 + NEX && !(NX0 > 0 && nt == NX0) && kvx) m = fmax(m, ex[nt - NX0][i]);
                }
                m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 4));
                m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 8));
                m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 16));
                if (g == 0) sm.red[warp][nt * 8 + 2 * t + i] = m;
            }
        }
        named_bar_sync(1, T_THREADS);
        if (threadIdx.x < T_WP) {
            double m = sm.red[0][threadIdx.x];
#pragma unroll
            for (int w = 1; w < T_WARPS; ++w) m = fmax(m, sm.red[w][threadIdx.x]);
            sm.xmax[0][threadIdx.x] = m;
            st_peer_f64(&sm.xmax[1][threadIdx.x], peer, m);
        }
        cluster_arrive();
        cluster_wait();
        // (2) exponentials and per-boot sums
#pragma unroll
        for (int nt = 0; nt < M_NT; ++nt) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int b = nt * 8 + 2 * t + i;
                const double M = fmax(sm.xmax[0][b], sm.xmax[1][b]);
                double e0 = kv0 ? exp(acc[0][nt][i] - M) : 0.0;
                double e1 = kv1 ? exp(acc[1][nt][i] - M) : 0.0;
                acc[0][nt][i] = e0;
                acc[1][nt][i] = e1;
                double s = e0 + e1;
                if constexpr (NEX > 0) {
                    if (nt >= NX0 && nt < NX0 + NEX) {
                        const bool ok = kvx && !(NX0 > 0 && nt == NX0);
                        double e2 = ok ? exp(ex[nt - NX0][i] - M) : 0.0;
                        ex[nt - NX0][i] = e2;
                        s += e2;
                    }
                }
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                if (g == 0) sm.red[warp][b] = s;
            }
        }
        named_bar_sync(1, T_THREADS);
        if (threadIdx.x < T_WP) {
            double s = sm.red[0][threadIdx.x];
#pragma unroll
            for (int w = 1; w < T_WARPS; ++w) s += sm.red[w][threadIdx.x];
            sm.xsum[0][threadIdx.x] = s;
            st_peer_f64(&sm.xsum[1][threadIdx.x], peer, s);
        }
        cluster_arrive();
        cluster_wait();
        // (3) jp[g, k] += sum_b e[k, b] / (S_b * scale)
        double r0 = 0.0, r1 = 0.0, rx = 0.0;
#pragma unroll
        for (int nt = 0; nt < M_NT; ++nt) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int b = nt * 8 + 2 * t + i;
                if (b < p.n_boot_pass) {
                    const double s0 = rank == 0 ? sm.xsum[0][b] : sm.xsum[1][b];
                    const double s1 = rank == 0 ? sm.xsum[1][b] : sm.xsum[0][b];
                    const double den = (s0 + s1) * p.scale;
                    r0 += acc[0][nt][i] / den;
                    r1 += acc[1][nt][i] / den;
                    if constexpr (NEX > 0) {
                        if (nt >= NX0 && nt < NX0 + NEX) rx += ex[nt - NX0][i] / den;
                    }
                }
            }
        }
        r0 += __shfl_xor_sync(0xffffffffu, r0, 1);
        r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
        r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
        r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
        if constexpr (NEX > 0) {
            rx += __shfl_xor_sync(0xffffffffu, rx, 1);
            rx += __shfl_xor_sync(0xffffffffu, rx, 2);
        }
        if (t == 0) {
            // jp is zero-filled by the caller; the shared grid tile receives two partial sums (commutative, so the
            // result does not depend on their order)
            double *out = p.jp + gene * p.ld_jp + kbase;
            if (kv0) atomicAdd(out + (m0 + 0) * 8 + g, r0);
            if (kv1) atomicAdd(out + (m0 + 1) * 8 + g, r1);
            if constexpr (NEX > 0) {
                if (kvx) atomicAdd(out + mx * 8 + g, rx);
            }
        }
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T_THREADS, 1) contract_mma_kernel(const TiledParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MmaSmem &sm = *reinterpret_cast<MmaSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t cid = cluster_id_x(), ncl = n_clusters_x();
    if (threadIdx.x == 0) {
        for (int s = 0; s < T_NS; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], T_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_arrive();
    cluster_wait();
    const int n_my_genes = ((int)cid < p.n_genes) ? (p.n_genes - (int)cid + (int)ncl - 1) / (int)ncl : 0;
    const int spg = (p.n_list + T_S - 1) / T_S;
    const int smsp = warp & 3, slot = warp >> 2;
    if (slot < 2)
        run_mma<0, 0>(p, sm, warp, lane, rank, n_my_genes, spg);
    else if ((smsp & 1) == 0)
        run_mma<7, 0>(p, sm, warp, lane, rank, n_my_genes, spg);
    else
        run_mma<7, 6>(p, sm, warp, lane, rank, n_my_genes, spg);
    cluster_arrive();
    cluster_wait();
}

// ------------------------------------------------------------------------------------------------
// generic kernel: one CTA per gene, threads stride the grid, boots in chunks of G_BC accumulators.
// Any K, any B.  When the grid fits one sweep (K <= 512) T is computed once per boot chunk; for larger grids the
// max / sum / accumulate phases each recompute it (this kernel is the fallback and the on-device cross-check,
// not the hot path).
constexpr int G_THREADS = 256;
constexpr int G_BC = 8;   // boots per chunk (divides 104, so a chunk never straddles two W passes)
constexpr int G_KPT = 2;  // grid points per thread per sweep -> 512 grid points per sweep

struct GenericParams {
    const double *table;
    int64_t ld_table;
    const int32_t *ridx;
    int64_t ld_ridx;
    const int32_t *cell_ids;
    int n_list;
    const double *W;  // pass-major [pass][n_w_rows][104]
    int64_t n_w_rows;
    int n_boot;
    double scale;
    int K;
    double *jp;
    int64_t ld_jp;
};

__device__ __forceinline__ void generic_accumulate(const GenericParams &p, int64_t g, int b0, int k0,
                                                   double (&acc)[G_KPT][G_BC]) {
#pragma unroll
    for (int i = 0; i < G_KPT; ++i)
#pragma unroll
        for (int j = 0; j < G_BC; ++j) acc[i][j] = 0.0;
    const double *Wc = p.W + ((int64_t)(b0 / WP_TILED) * p.n_w_rows) * WP_TILED + (b0 % WP_TILED);
    for (int c = 0; c < p.n_list; ++c) {
        const int col = p.cell_ids ? p.cell_ids[c] : c;
        const int64_t row = p.ridx[g * p.ld_ridx + col];
        const double *a = p.table + row * p.ld_table;
        double av[G_KPT];
#pragma unroll
        for (int i = 0; i < G_KPT; ++i) {
            int k = k0 + i * G_THREADS;
            av[i] = k < p.K ? a[k] : 0.0;
        }
        const double *w = Wc + (int64_t)c * WP_TILED;
#pragma unroll
        for (int j = 0; j < G_BC; ++j) {
            double wv = (b0 + j < p.n_boot) ? w[j] : 0.0;
#pragma unroll
            for (int i = 0; i < G_KPT; ++i) acc[i][j] = fma(av[i], wv, acc[i][j]);
        }
    }
}

__global__ void __launch_bounds__(G_THREADS) contract_generic_kernel(const GenericParams p) {
    extern __shared__ double gs[];  // [K] jp accumulator, then reduction scratch
    double *s_jp = gs;
    double *s_red = gs + p.K;  // [G_THREADS/32][G_BC]
    __shared__ double s_m[G_BC], s_s[G_BC];
    const int64_t g = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = p.K;
    for (int k = threadIdx.x; k < K; k += G_THREADS) s_jp[k] = 0.0;
    __syncthreads();
    const int sweeps = (K + G_THREADS * G_KPT - 1) / (G_THREADS * G_KPT);
    for (int b0 = 0; b0 < p.n_boot; b0 += G_BC) {
        double mloc[G_BC], sloc[G_BC];
#pragma unroll
        for (int j = 0; j < G_BC; ++j) {
            mloc[j] = -INFINITY;
            sloc[j] = 0;
        }
        double acc[G_KPT][G_BC];
        if (sweeps == 1) generic_accumulate(p, g, b0, threadIdx.x, acc);
        for (int phase = 0; phase < 3; ++phase) {
            for (int sw = 0; sw < sweeps; ++sw) {
                const int k0 = sw * G_THREADS * G_KPT + threadIdx.x;
                if (sweeps > 1) generic_accumulate(p, g, b0, k0, acc);
                if (phase == 0) {
#pragma unroll
                    for (int j = 0; j < G_BC; ++j)
#pragma unroll
                        for (int i = 0; i < G_KPT; ++i)
                            if (k0 + i * G_THREADS < K) mloc[j] = fmax(mloc[j], acc[i][j]);
                } else if (phase == 1) {
#pragma unroll
                    for (int j = 0; j < G_BC; ++j)
#pragma unroll
                        for (int i = 0; i < G_KPT; ++i)
                            if (k0 + i * G_THREADS < K) sloc[j] += exp(acc[i][j] - s_m[j]);
                } else {
#pragma unroll
                    for (int i = 0; i < G_KPT; ++i) {
                        int k = k0 + i * G_THREADS;
                        if (k < K) {
                            double r = 0;
#pragma unroll
                            for (int j = 0; j < G_BC; ++j)
                                if (b0 + j < p.n_boot) r += exp(acc[i][j] - s_m[j]) / (s_s[j] * p.scale);
                            s_jp[k] += r;
                        }
                    }
                }
            }
            if (phase == 0) {
#pragma unroll
                for (int j = 0; j < G_BC; ++j) {
                    double m = mloc[j];
                    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
                    if (lane == 0) s_red[warp * G_BC + j] = m;
                }
                __syncthreads();
                if (threadIdx.x < G_BC) {
                    double m = s_red[threadIdx.x];
                    for (int w = 1; w < G_THREADS / 32; ++w) m = fmax(m, s_red[w * G_BC + threadIdx.x]);
                    s_m[threadIdx.x] = m;
                }
                __syncthreads();
            } else if (phase == 1) {
#pragma unroll
                for (int j = 0; j < G_BC; ++j) {
                    double s = sloc[j];
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (lane == 0) s_red[warp * G_BC + j] = s;
                }
                __syncthreads();
                if (threadIdx.x < G_BC) {
                    double s = s_red[threadIdx.x];
                    for (int w = 1; w < G_THREADS / 32; ++w) s += s_red[w * G_BC + threadIdx.x];
                    s_s[threadIdx.x] = s;
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k < K; k += G_THREADS) p.jp[g * p.ld_jp + k] = s_jp[k];
}

// ------------------------------------------------------------------------------------------------
// ensemble form: jp[g, :] = normalise( sum_c normalise_k(exp(lp[c, x[g,c], :])) )
__global__ void row_expsum_kernel(const double *__restrict__ table, int64_t ld_table, int K, int64_t n_rows,
                                  double *__restrict__ rowsum) {
    int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    int lane = threadIdx.x & 31;
    double s = 0;
    for (int k = lane; k < K; k += 32) s += exp(table[row * ld_table + k]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) rowsum[row] = s;
}

__global__ void ensemble_kernel(const double *__restrict__ table, int64_t ld_table, const int32_t *__restrict__ ridx,
                                int64_t ld_ridx, const int32_t *__restrict__ cell_ids, int n_list, int K,
                                const double *__restrict__ rowsum, double *__restrict__ jp, int64_t ld_jp) {
    extern __shared__ double es[];  // [K]
    __shared__ double s_red[32];
    const int64_t g = blockIdx.x;
    double tot = 0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double acc = 0;
        for (int c = 0; c < n_list; ++c) {
            int col = cell_ids ? cell_ids[c] : c;
            int64_t row = ridx[g * ld_ridx + col];
            acc += exp(table[row * ld_table + k]) / rowsum[row];
        }
        es[k] = acc;
        tot += acc;
    }
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) s_red[0] = v;
    }
    __syncthreads();
    const double s = s_red[0];
    for (int k = threadIdx.x; k < K; k += blockDim.x) jp[g * ld_jp + k] = es[k] / s;
}

// ------------------------------------------------------------------------------------------------
__global__ void gather_modes_kernel(const int32_t *__restrict__ ridx, int64_t ld_ridx, int G, int n_cells,
                                    const int32_t *__restrict__ row_mode, const double *__restrict__ mag,
                                    double *__restrict__ modes) {
    // modes is G x n_cells column-major; ridx is gene-major: transpose through a tile
    __shared__ double tile[32][33];
    int gx = blockIdx.x * 32, cy = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int g = gx + j, c = cy + threadIdx.x;
        if (g < G && c < n_cells) tile[j][threadIdx.x] = mag[row_mode[ridx[(int64_t)g * ld_ridx + c]]];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int c = cy + j, g = gx + threadIdx.x;
        if (g < G && c < n_cells) modes[(int64_t)c * G + g] = tile[threadIdx.x][j];
    }
}

__global__ void gather_post_kernel(const int32_t *__restrict__ ridx, int64_t ld_ridx, int G, int n_cells,
                                   const double *__restrict__ table, int64_t ld_table, int K, double sentinel,
                                   double minlogprob, double *__restrict__ post) {
    // post[c][g + G*k]; one CTA per (gene tile of 32, cell), tile-transposed so both sides are coalesced
    __shared__ double tile[32][33];
    const int c = blockIdx.y, gx = blockIdx.x * 32;
    for (int k0 = 0; k0 < K; k0 += 32) {
        for (int j = threadIdx.y; j < 32; j += blockDim.y) {
            int g = gx + j, k = k0 + threadIdx.x;
            if (g < G && k < K) {
                double v = table[(int64_t)ridx[(int64_t)g * ld_ridx + c] * ld_table + k];
                tile[j][threadIdx.x] = (v <= sentinel) ? minlogprob : v;
            }
        }
        __syncthreads();
        for (int j = threadIdx.y; j < 32; j += blockDim.y) {
            int k = k0 + j, g = gx + threadIdx.x;
            if (g < G && k < K) post[((int64_t)c * K + k) * G + g] = tile[threadIdx.x][j];
        }
        __syncthreads();
    }
}

__global__ void transpose_out_kernel(const double *__restrict__ src, int64_t ld_src, int G, int K,
                                     double *__restrict__ dst) {
    // src [G][ld_src] gene-major -> dst G x K column-major
    __shared__ double tile[32][33];
    int gx = blockIdx.x * 32, ky = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int g = gx + j, k = ky + threadIdx.x;
        if (g < G && k < K) tile[j][threadIdx.x] = src[(int64_t)g * ld_src + k];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int k = ky + j, g = gx + threadIdx.x;
        if (g < G && k < K) dst[(int64_t)k * G + g] = tile[threadIdx.x][j];
    }
}

__global__ void transpose_in_kernel(const double *__restrict__ src, int G, int K, double *__restrict__ dst,
                                    int64_t ld_dst) {
    // src G x K column-major -> dst [G][ld_dst] gene-major (columns K..ld_dst untouched)
    __shared__ double tile[32][33];
    int gx = blockIdx.x * 32, ky = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int k = ky + j, g = gx + threadIdx.x;
        if (g < G && k < K) tile[j][threadIdx.x] = src[(int64_t)k * G + g];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int g = gx + j, k = ky + threadIdx.x;
        if (g < G && k < K) dst[(int64_t)g * ld_dst + k] = tile[threadIdx.x][j];
    }
}

// register-resident DFMA loop: 16 independent chains per thread
__global__ void fp64_peak_kernel(double *sink, int iters) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const double x = 1.0000000001, y = 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
}

}  // namespace

cudaError_t launch_build_w(const int32_t *boot_idx, int n_boot, int D, int n_list, double *W, int n_w_rows,
                           cudaStream_t st) {
    const int passes = (n_boot + WP_TILED - 1) / WP_TILED;
    cudaError_t e = cudaMemsetAsync(W, 0, sizeof(double) * (size_t)(passes > 0 ? passes : 1) * n_w_rows * WP_TILED, st);
    if (e != cudaSuccess) return e;
    if (n_boot <= 0 || D <= 0) return cudaSuccess;
    build_w_kernel<<<n_boot, 256, 0, st>>>(boot_idx, n_boot, D, n_list, W, n_w_rows);
    return cudaGetLastError();
}

bool contract_tiled_supported(const ContractArgs &a) {
    return a.K <= KP_TILED && a.ld_table == KP_TILED && a.n_list >= 1 && a.n_boot >= 1 &&
           a.n_w_rows >= round_up(a.n_list, 8);
}

cudaError_t launch_contract_tiled(const ContractArgs &a, int n_sm, int variant, cudaStream_t st, int *n_launches) {
    if (a.n_genes <= 0) return cudaSuccess;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(contract_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(TiledSmem));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(contract_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MmaSmem));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    int clusters = n_sm / 2;
    if (clusters < 1) clusters = 1;
    if (clusters > a.n_genes) clusters = a.n_genes;
    const int passes = (a.n_boot + WP_TILED - 1) / WP_TILED;
    for (int ps = 0; ps < passes; ++ps) {
        TiledParams p;
        p.table = a.table;
        p.ridx = a.ridx;
        p.ld_ridx = a.ld_ridx;
        p.cell_ids = a.cell_ids;
        p.n_list = a.n_list;
        p.W = a.W + (size_t)ps * a.n_w_rows * WP_TILED;
        p.ldw = WP_TILED;
        p.n_boot_pass = (a.n_boot - ps * WP_TILED) < WP_TILED ? (a.n_boot - ps * WP_TILED) : WP_TILED;
        p.scale = a.scale;
        p.n_genes = a.n_genes;
        p.K = a.K;
        p.jp = a.jp;
        p.ld_jp = a.ld_jp;
        p.accumulate = ps > 0;
        if (variant == 1)
            contract_tiled_kernel<<<2 * clusters, T_THREADS, sizeof(TiledSmem), st>>>(p);
        else
            contract_mma_kernel<<<2 * clusters, T_THREADS, sizeof(MmaSmem), st>>>(p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (n_launches) ++*n_launches;
    }
    return cudaSuccess;
}

cudaError_t launch_contract_generic(const ContractArgs &a, cudaStream_t st, int *n_launches) {
    if (a.n_genes <= 0) return cudaSuccess;
    size_t smem = sizeof(double) * ((size_t)a.K + (G_THREADS / 32) * G_BC);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(contract_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    GenericParams p;
    p.table = a.table;
    p.ld_table = a.ld_table;
    p.ridx = a.ridx;
    p.ld_ridx = a.ld_ridx;
    p.cell_ids = a.cell_ids;
    p.n_list = a.n_list;
    p.W = a.W;
    p.n_w_rows = a.n_w_rows;
    p.n_boot = a.n_boot;
    p.scale = a.scale;
    p.K = a.K;
    p.jp = a.jp;
    p.ld_jp = a.ld_jp;
    contract_generic_kernel<<<a.n_genes, G_THREADS, smem, st>>>(p);
    if (n_launches) ++*n_launches;
    return cudaGetLastError();
}

cudaError_t launch_ensemble(const ContractArgs &a, double *rownorm_scratch, int64_t n_rows, cudaStream_t st) {
    if (a.n_genes <= 0) return cudaSuccess;
    int wpb = 8;
    row_expsum_kernel<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(a.table, a.ld_table, a.K, n_rows,
                                                                               rownorm_scratch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    size_t smem = sizeof(double) * a.K;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(ensemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    ensemble_kernel<<<a.n_genes, 256, smem, st>>>(a.table, a.ld_table, a.ridx, a.ld_ridx, a.cell_ids, a.n_list, a.K,
                                                  rownorm_scratch, a.jp, a.ld_jp);
    return cudaGetLastError();
}

cudaError_t launch_gather_modes(const int32_t *ridx, int ld_ridx, int G, int n_cells, const int32_t *row_mode,
                                const double *mag, double *modes, cudaStream_t st) {
    if (G <= 0 || n_cells <= 0) return cudaSuccess;
    dim3 grid((G + 31) / 32, (n_cells + 31) / 32), block(32, 8);
    gather_modes_kernel<<<grid, block, 0, st>>>(ridx, ld_ridx, G, n_cells, row_mode, mag, modes);
    return cudaGetLastError();
}

cudaError_t launch_gather_post(const int32_t *ridx, int ld_ridx, int G, int n_cells, const double *table,
                               int ld_table, int K, double sentinel, double minlogprob, double *post,
                               cudaStream_t st) {
    if (G <= 0 || n_cells <= 0) return cudaSuccess;
    dim3 grid((G + 31) / 32, n_cells), block(32, 8);
    gather_post_kernel<<<grid, block, 0, st>>>(ridx, ld_ridx, G, n_cells, table, ld_table, K, sentinel, minlogprob, post);
    return cudaGetLastError();
}

cudaError_t launch_transpose_out(const double *src, int ld_src, int G, int K, double *dst, cudaStream_t st) {
    if (G <= 0 || K <= 0) return cudaSuccess;
    dim3 grid((G + 31) / 32, (K + 31) / 32), block(32, 8);
    transpose_out_kernel<<<grid, block, 0, st>>>(src, ld_src, G, K, dst);
    return cudaGetLastError();
}

cudaError_t launch_transpose_in(const double *src, int G, int K, double *dst, int ld_dst, cudaStream_t st) {
    if (G <= 0 || K <= 0) return cudaSuccess;
    dim3 grid((G + 31) / 32, (K + 31) / 32), block(32, 8);
    transpose_in_kernel<<<grid, block, 0, st>>>(src, G, K, dst, ld_dst);
    return cudaGetLastError();
}

cudaError_t launch_fp64_peak(double *sink, int iters, int blocks, cudaStream_t st) {
    fp64_peak_kernel<<<blocks, 256, 0, st>>>(sink, iters);
    return cudaGetLastError();
}

}  // namespace scde
