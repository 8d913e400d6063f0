// contract_i8.cu -- the bootstrap joint posterior (src/jpmatLogBoot.cpp:216-275, :460-499) on the 5th-generation tensor
// cores: tcgen05.mma.kind::i8 with 32-bit integer accumulators in tensor memory.
//
// Per gene the reference computes  T[b, k] = sum_c W[c, b] * lp[c, x[g, c], k]  (W = how often boot b drew cell c) and
// then jp[g, :] = (1/B) sum_b softmax_k T[b, :].  boot_contract.cu evaluates that sum on the FP64 pipe (DMMA), where it
// is bound by the 37 TFLOP/s FP64 rate.  Here the table is held in FIXED POINT instead: in the zero-base form every
// table value D = lp(x) - lp(0) (or lp(x) itself) lies in [-751, 751] -- log of a normalised double -- or is the
// reference's "log 0" sentinel (-DBL_MAX/n/1.1, :127,204).  D is rounded to a multiple of 2^-29 and split into five
// signed radix-256 digits (int8 planes 0..4, value = 2^-29 * sum_p 256^p d_p).  W is a small non-negative integer
// (<= 127 is checked).  Every product and every sum is then EXACT in int32 (|sum| <= 128 * draws), the planes are
// recombined in int64 and converted to FP64 once per (boot, grid point), so the only error is the table rounding:
// |T error| <= draws * 2^-30 in the worst case and about sqrt(2 * draws) * 2^-29 / sqrt(12) typically (5e-8 at 5000
// cells per group) -- inside the 1e-6 tolerance on log-posteriors, and independent of the summation order, so the
// result is deterministic by construction.
//
// The sentinel is not summed at all.  A grid point of (gene, boot) is "log 0" exactly when some drawn row is "log 0"
// there; the non-sentinel points of a row form one interval [klo, khi] (common.cuh), so it is enough to intersect the
// intervals of the drawn rows: sentinel_range_kernel does that per (gene, boot) with packed 16-bit max / min, and the
// epilogue writes the sentinel outside the intersection.  (gene, boot) pairs with an EMPTY intersection -- where the
// reference's soft-max would compare multiples of the sentinel -- and irregular rows raise flag 4: the caller reruns on
// the FP64 kernel, which adds the sentinels as the reference does.
//
// Kernel: persistent, one CTA per SM, three warp roles, no CTA-wide barrier in the steady state.
//   * work item = (gene, piece): one 512-byte piece of the gene's table rows = 102 grid points x 5 planes.  The item
//     accumulates D[boot (128 lanes, 104 real)][plane * 102 + i] in all 512 tensor-memory columns of the SM.
//   * producers (Q_PGROUPS x 4 warps; Q_PGROUPS = 1, see launch_contract_i8_pass): gather the list entries' pieces (512 contiguous, 512-byte-aligned bytes per
//     entry) and W rows (128 bytes) with 16-byte cp.async straight into the canonical 128-byte-swizzle layout of an
//     MN-major operand; completion is signalled by cp.async.mbarrier.arrive; a 10-stage ring of 32 entries (20 KB per
//     stage) keeps up to 200 KB in flight per SM.  The loop is branch-free per stage: constant offsets, slot / phase
//     counters instead of divisions, list entries fetched 8-16 of the group's stages ahead.
//   * MMA (1 thread): per stage two tcgen05.mma (M = 128 boots, N = 256, K = 32 entries), then tcgen05.commit onto the
//     stage's "empty" barrier; after the last stage a commit onto the accumulator barrier.
//   * epilogue (4 warps, one per 32-lane quarter of tensor memory): tcgen05.ld the five planes of 8 grid points at a
//     time, recombine into one 64-bit integer per (boot, grid point) and store it; softmax_i8_warp_kernel finishes the gene
//     (conversion, zero-count base Z[b, k], sentinel ranges, soft-max, average).  Producers keep prefetching the next
//     item meanwhile.
// Roofline: HBM gather bandwidth (512 bytes per visited (gene, cell) pair and piece; tools/gather_bench.cu measures
// 6.8 TB/s for this access pattern with 128 KB in flight per SM); the tensor pipe needs 2 x 128 cycles per 32 entries.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "fastmath.cuh"
#include <cfloat>
#include <cmath>
#include <cstdlib>

namespace scde {
namespace {

using namespace ptx;

constexpr int Q_ES = 32;                         // list entries per stage = K of one tcgen05.mma.kind::i8
constexpr int Q_NS = 10;                         // ring depth (largest; the kernel takes the depth as a template parameter)
constexpr int Q_A_BYTES = Q_ES * Q_WB;           // W tile of a stage: 4096
constexpr int Q_B_BYTES = Q_ES * Q_PIECE;        // table tile of a stage: 16384
constexpr int Q_STAGE_BYTES = Q_A_BYTES + Q_B_BYTES;  // 20480, a multiple of the 1024-byte swizzle atom
// PG producer warp groups (template parameter of the kernel); group g gathers the stages s = g (mod PG) of an item.
// Eight epilogue warps: a warp may only read the tensor-memory lanes 32 (warp id mod 4) .. + 31, so two warps share each
// quarter and split its columns.  (With four warps the epilogue of an item took longer than the ring can cover, and the
// tensor memory is not free for the next item before it ends.)
constexpr int Q_EPILOGUE_WARPS = 8;
constexpr int q_threads(int pg) { return (4 * pg + Q_EPILOGUE_WARPS + 1) * 32; }  // + the MMA warp
constexpr int Q_EPI_ROUNDS = (Q_PW + 7) / 8;  // rounds of 8 grid points per piece
constexpr int Q_TMEM_COLS = 512;
constexpr long long Q_WATCHDOG_CYCLES = 4000000000ll;  // a barrier wait longer than ~2 s aborts the kernel (err = 2)
static_assert(Q_NV * Q_PW <= Q_TMEM_COLS && Q_NV * Q_PW <= Q_PIECE, "a piece is one pass of tensor memory");
static_assert((Q_PW & 1) == 0, "the epilogue reads pairs of columns");

struct I8Smem {
    uint64_t full[Q_NS];
    uint64_t empty[Q_NS];
    uint64_t acc_full, acc_empty;
    uint32_t tmem_base;
    volatile int abort;
};
constexpr int Q_XBUF_BYTES = Q_EPILOGUE_WARPS * 2048;  // per epilogue warp: one [32 boots][8 grid points] tile of 64-bit sums
constexpr size_t q_smem_bytes(int ns) { return 1024 /* alignment slack */ + (size_t)ns * Q_STAGE_BYTES + Q_XBUF_BYTES + sizeof(I8Smem); }

struct I8Params {
    const int8_t *qtable;  // [rows][ldq]: row = [piece][plane][102] (+ 2)
    int64_t ldq;
    const int32_t *lst_row, *lst_cell, *lst_len, *order;
    int64_t ld_lst;
    const int8_t *W8;    // this pass: [n_w_rows][128]
    long long *T;        // [n_pos][104][416]: 2^29 * T[boot, grid] as exact integers (without Z, without sentinels)
    int n_boot;          // real boots of this pass: rows beyond are not stored
    int n_pos;           // genes in this launch: positions [0, n_pos) of `order`
    int n_pieces;        // pieces per gene
    int32_t *err;        // device flag: 2 = watchdog abort
    unsigned long long *dbg;  // optional diagnostics: [0] += cycles between "accumulators ready" and "tensor memory released",
                              // [1] += items, [2] += cycles the MMA thread waited for the release (epilogue warp 0 / MMA thread)
    const int8_t *W8b;   // twin launch: the second joint's W (same cells, same lists, other draws), else NULL
    long long *Tb;       // twin launch: the second joint's T tiles
    int twin;            // 1: items come in pairs (2 i, 2 i + 1) = the same (gene, piece) for joint 0 and joint 1, so that two
                         // neighbouring CTAs gather the same table rows at the same time and the second read is an L2 hit
    int piece_major;     // item order: 0 = (gene, piece) with the piece fastest, 1 = (piece, gene) with the gene fastest
    int cold_evict_first;  // HINT kernels: rows without the hot bit are loaded evict_first (else without a hint)
};

__device__ __forceinline__ bool wait_or_abort(I8Smem &sm, uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (sm.abort) return false;
        if (clock64() - t0 > Q_WATCHDOG_CYCLES) {  // never in a correct run: leave instead of hanging the GPU
            sm.abort = 1;
            return false;
        }
    }
    return true;
}

// item -> (gene position, piece).  Gene-major order: the pieces of one gene go to neighbouring CTAs of the same round, so
// the W rows and list entries they share are L2 hits for three of the four.  Piece-major order: all CTAs work on the same
// 512-byte piece of the table rows at the same time -- four times as many genes in flight per piece, and a quarter of
// the table as the working set of a phase, so the rows of small counts (shared by thousands of genes) are L2 hits more
// often; the lists and W rows are re-read once per phase (8 + 128 B per entry against the piece's 512 B).
struct Item {
    int pos, piece;
};
__device__ __forceinline__ Item decode_item(const I8Params &p, int item) {
    Item it;
    item >>= p.twin;  // twin launch: bit 0 of the item selects the joint
    if (p.piece_major) {
        it.piece = item / p.n_pos;
        it.pos = item - it.piece * p.n_pos;
    } else {
        it.pos = item / p.n_pieces;
        it.piece = item - it.pos * p.n_pieces;
    }
    return it;
}
__device__ __forceinline__ int64_t item_gene(const I8Params &p, int item) {
    item >>= p.twin;
    const int pos = p.piece_major ? item % p.n_pos : item / p.n_pieces;
    return p.order ? p.order[pos] : pos;
}

// ---- producers ------------------------------------------------------------------------------------------------------
// Warp (grp, kg) gathers entries 8 kg .. 8 kg + 7 (one k-group of the MMA) of the stages s = grp (mod Q_PGROUPS) of
// every item.  Canonical 128-byte-swizzle layout of an MN-major operand: the bytes of one entry are split into runs of
// 128 (8 pieces of 16 bytes); run r of entry kk of k-group kg sits at [r][kg][kk][128 B] and its piece j at
// 16 * (j ^ kk).  Lane (l3 = lane >> 3, l7 = lane & 7) copies piece l7 of every run of entries l3 and l3 + 4: a warp
// instruction moves four 128-byte runs of global memory into four 128-byte lines of shared memory -- coalesced on both
// sides, no bank conflicts.
//
// List entries: lane (l3, l7) holds entries l3 and l3 + 4 of the warp's own-stage 8 t + l7 -- one load instruction
// fetches the entries of eight own-stages, issued one block (8 own-stages = 8 Q_PGROUPS stages) before its first use, the
// first two blocks at the start of the item; the identity of the NEXT item (gene, list length) is loaded one item
// ahead.  An earlier version with a shorter look-ahead spent 17 % of the producers' cycles waiting on these loads and
// 55 % issuing ~180 dependent instructions per stage (profiles/r01v); this loop issues about 45.
template <int DOFF, int SOFF>
__device__ __forceinline__ void cp_async16_at(uint32_t dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0+%2], [%1+%3], 16;" ::"r"(dst_smem), "l"(src), "n"(DOFF), "n"(SOFF) : "memory");
}
template <int DOFF, int SOFF>
__device__ __forceinline__ void cp_async16_hint_at(uint32_t dst_smem, const void *src, uint64_t policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0+%2], [%1+%3], 16, %4;" ::"r"(dst_smem), "l"(src), "n"(DOFF), "n"(SOFF),
                 "l"(policy)
                 : "memory");
}
// HINT: list entries carry LIST_HOT_BIT in `cell` when the row is one of the first few of its cell (a small count: such a
// row is shared by thousands of genes); its pieces are loaded with the L2 evict_last policy, the others evict_first (or
// unhinted), so that the stream of rows that are used once does not push the shared ones out of the L2.
template <int Q_PGROUPS, bool HINT, int NS>
__device__ __forceinline__ void run_producer(const I8Params &p, I8Smem &sm, uint32_t stage0, int n_items, int warp, int lane) {
    uint64_t pol_hot = 0, pol_cold = 0;
    if (HINT) {
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_hot));
        if (p.cold_evict_first)
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_cold));
        else
            asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_cold));
    }
    const int grp = warp >> 2, kg = warp & 3;
    const int l7 = lane & 7, l3 = lane >> 3;
    // entry kk = l3 (and l3 + 4): line kk of this warp's k-group, piece l7 at 16 * (l7 ^ kk)
    const uint32_t d0 = (uint32_t)kg * 1024u + (uint32_t)l3 * 128u + (uint32_t)((l7 ^ l3) << 4);
    const uint32_t d1 = (uint32_t)kg * 1024u + (uint32_t)(l3 + 4) * 128u + (uint32_t)((l7 ^ (l3 + 4)) << 4);
    int slot_b = 0;        // ring slot of the first stage of the current item
    uint32_t fill_b = 0;   // how often that slot has been filled before
    int item = blockIdx.x;
    int64_t gene_n = 0;
    int len_n = 0;
    if (item < n_items) {
        gene_n = item_gene(p, item);
        len_n = p.lst_len[gene_n];
    }
    for (; item < n_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        const int64_t gene = gene_n;
        const int nst = (len_n + Q_ES - 1) / Q_ES;
        if (item + (int)gridDim.x < n_items) {
            gene_n = item_gene(p, item + gridDim.x);
            len_n = p.lst_len[gene_n];
        }
        const int8_t *qbase = p.qtable + (int64_t)it.piece * Q_PIECE + l7 * 16;
        const int8_t *wbase = ((p.twin && (item & 1)) ? p.W8b : p.W8) + l7 * 16;
        const int32_t *lrow = p.lst_row + gene * p.ld_lst + kg * 8 + l3;
        const int32_t *lcell = p.lst_cell + gene * p.ld_lst + kg * 8 + l3;
        int32_t r0a = 0, r1a = 0, c0a = 0, c1a = 0, r0b = 0, r1b = 0, c0b = 0, c1b = 0;
        auto load_block = [&](int t, int32_t &r0, int32_t &r1, int32_t &c0, int32_t &c1) {
            const int s = grp + Q_PGROUPS * (8 * t + l7);
            r0 = r1 = c0 = c1 = 0;
            if (s < nst) {
                r0 = __ldg(lrow + s * Q_ES);
                r1 = __ldg(lrow + s * Q_ES + 4);
                c0 = __ldg(lcell + s * Q_ES);
                c1 = __ldg(lcell + s * Q_ES + 4);
            }
        };
        load_block(0, r0a, r1a, c0a, c1a);
        load_block(1, r0b, r1b, c0b, c1b);
        int slot = slot_b + grp;
        uint32_t fill = fill_b;
        if (slot >= NS) {
            slot -= NS;
            ++fill;
        }
        for (int t = 0; grp + Q_PGROUPS * 8 * t < nst; ++t) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int s = grp + Q_PGROUPS * (8 * t + u);
                if (s < nst) {  // warp-uniform
                    const int from = (lane & 24) | u;
                    const int32_t row0 = __shfl_sync(0xffffffffu, r0a, from), row1 = __shfl_sync(0xffffffffu, r1a, from);
                    const int32_t cel0 = __shfl_sync(0xffffffffu, c0a, from), cel1 = __shfl_sync(0xffffffffu, c1a, from);
                    const uint32_t sA = stage0 + (uint32_t)slot * Q_STAGE_BYTES;
                    const uint32_t sB0 = sA + Q_A_BYTES + d0, sB1 = sA + Q_A_BYTES + d1;
                    const int8_t *src0 = qbase + (int64_t)row0 * p.ldq;
                    const int8_t *src1 = qbase + (int64_t)row1 * p.ldq;
                    const int8_t *w0 = wbase + (int64_t)(cel0 & ~LIST_HOT_BIT) * Q_WB;
                    const int8_t *w1 = wbase + (int64_t)(cel1 & ~LIST_HOT_BIT) * Q_WB;
                    if (fill > 0 && !wait_or_abort(sm, &sm.empty[slot], (fill - 1) & 1u)) return;
                    if (HINT) {
                        const uint64_t pol0 = (cel0 & LIST_HOT_BIT) ? pol_hot : pol_cold;
                        const uint64_t pol1 = (cel1 & LIST_HOT_BIT) ? pol_hot : pol_cold;
                        cp_async16_hint_at<0, 0>(sB0, src0, pol0);
                        cp_async16_hint_at<0, 0>(sB1, src1, pol1);
                        cp_async16_hint_at<4096, 128>(sB0, src0, pol0);
                        cp_async16_hint_at<4096, 128>(sB1, src1, pol1);
                        cp_async16_hint_at<8192, 256>(sB0, src0, pol0);
                        cp_async16_hint_at<8192, 256>(sB1, src1, pol1);
                        cp_async16_hint_at<12288, 384>(sB0, src0, pol0);
                        cp_async16_hint_at<12288, 384>(sB1, src1, pol1);
                        cp_async16_hint_at<0, 0>(sA + d0, w0, pol_hot);
                        cp_async16_hint_at<0, 0>(sA + d1, w1, pol_hot);
                    } else {
                        cp_async16_at<0, 0>(sB0, src0);
                        cp_async16_at<0, 0>(sB1, src1);
                        cp_async16_at<4096, 128>(sB0, src0);
                        cp_async16_at<4096, 128>(sB1, src1);
                        cp_async16_at<8192, 256>(sB0, src0);
                        cp_async16_at<8192, 256>(sB1, src1);
                        cp_async16_at<12288, 384>(sB0, src0);
                        cp_async16_at<12288, 384>(sB1, src1);
                        cp_async16_at<0, 0>(sA + d0, w0);
                        cp_async16_at<0, 0>(sA + d1, w1);
                    }
                    cp_async_mbar_arrive_noinc(&sm.full[slot]);
                    slot += Q_PGROUPS;
                    if (slot >= NS) {
                        slot -= NS;
                        ++fill;
                    }
                }
            }
            r0a = r0b;
            r1a = r1b;
            c0a = c0b;
            c1a = c1b;
            load_block(t + 2, r0b, r1b, c0b, c1b);
        }
        slot_b += nst % NS;
        fill_b += (uint32_t)(nst / NS);
        if (slot_b >= NS) {
            slot_b -= NS;
            ++fill_b;
        }
    }
}

// ---- MMA issuer (one thread) ------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void run_mma(const I8Params &p, I8Smem &sm, uint32_t stage0, uint32_t tmem, int n_items) {
    const uint32_t idesc = umma_idesc_s8_mn(128, 256);
    int slot = 0;
    uint32_t fill = 0, n_done = 0;
    int item = blockIdx.x;
    int len_n = item < n_items ? p.lst_len[item_gene(p, item)] : 0;
    for (; item < n_items; item += gridDim.x, ++n_done) {
        const int nst = (len_n + Q_ES - 1) / Q_ES;
        if (item + (int)gridDim.x < n_items) len_n = p.lst_len[item_gene(p, item + gridDim.x)];
        if (n_done > 0) {  // the epilogue has drained the previous item's accumulators
            const long long t0 = p.dbg ? clock64() : 0;
            if (!wait_or_abort(sm, &sm.acc_empty, (n_done - 1) & 1u)) return;
            tc_fence_after_sync();
            if (p.dbg) atomicAdd(p.dbg + 2, (unsigned long long)(clock64() - t0));
        }
        for (int s = 0; s < nst; ++s) {
            if (!wait_or_abort(sm, &sm.full[slot], fill & 1u)) return;
            fence_proxy_async_smem();
            tc_fence_after_sync();
            const uint32_t sA = stage0 + (uint32_t)slot * Q_STAGE_BYTES;
            const uint32_t sB = sA + Q_A_BYTES;
            const uint64_t da = umma_desc_sw128(sA, 4096u, 1024u);
            umma_s8(tmem, da, umma_desc_sw128(sB, 4096u, 1024u), idesc, s > 0);                      // runs 0, 1
            umma_s8(tmem + 256u, da, umma_desc_sw128(sB + 2u * 4096u, 4096u, 1024u), idesc, s > 0);  // runs 2, 3
            umma_commit(&sm.empty[slot]);  // frees the slot once these MMAs have read it
            if (++slot == NS) {
                slot = 0;
                ++fill;
            }
        }
        if (nst > 0)
            umma_commit(&sm.acc_full);
        else
            mbar_arrive(&sm.acc_full);
    }
}

// ---- epilogue: warps e = 0..7; warp e reads tensor-memory lanes 32 (e & 3) .. + 31 and the rounds of half e >> 2 --------
// The tensor memory is not free for the next item's MMAs before the epilogue ends (all 512 columns belong to one item), so
// the epilogue does the minimum: tcgen05.ld 8 grid points x 5 planes, recombine the plane sums into ONE 64-bit integer
// per (boot, grid point) and store it with streaming stores -- no loads, no floating point.  |sum_p| <= 128 * draws < 2^23
// (draws <= 65000 is checked by the launcher), so a = s0 + 256 s1 and b = s2 + 256 s3 fit 32 bits and the value is
// a + 2^16 b + 2^32 s4.  Conversion to FP64, the zero-count base Z and the sentinel ranges are applied by
// softmax_i8_warp_kernel when it reads T.  (An epilogue that did all of that took 16 800 cycles per item -- 8.9 us, of which
// the ring hides 3 -- and cost a quarter of the kernel: SCDE_B200_EPI_TIMING, profiles/r01x.)
__device__ __forceinline__ void run_epilogue(const I8Params &p, I8Smem &sm, uint32_t xbuf, uint32_t tmem, int n_items, int ewarp,
                                             int lane) {
    const int qd = ewarp & 3, half = ewarp >> 2;
    const uint32_t tlane = tmem + ((uint32_t)(qd * 32) << 16);
    // A thread holds 8 grid points of ONE boot (64 bytes of a T row; rows are 3328 bytes apart): stored directly, a warp
    // store touches 32 lines.  The round's [32 boots][8 points] tile is turned through shared memory instead, so that a
    // store instruction writes the complete 64-byte runs of 8 boots (8 lines, every sector full).  Tile layout: 16-byte
    // chunk j of boot l at unit (l + 8 j) mod 32 of row j -- conflict-free both ways.
    const uint32_t xw = xbuf + (uint32_t)ewarp * 2048u;
    const int rb = lane >> 2, rc = lane & 3;  // read side: boot rb + 8 i of the warp, chunk rc
    const int r_begin = half ? (Q_EPI_ROUNDS + 1) / 2 : 0, r_end = half ? Q_EPI_ROUNDS : (Q_EPI_ROUNDS + 1) / 2;
    uint32_t n_done = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
        const Item it = decode_item(p, item);
        const int64_t gene = p.order ? p.order[it.pos] : it.pos;
        const bool empty_list = p.lst_len[gene] <= 0;
        while (!mbar_try_wait(&sm.acc_full, n_done & 1u)) {  // an item takes tens of microseconds: sleep, do not spin
            __nanosleep(200);
            if (sm.abort) return;
        }
        tc_fence_after_sync();
        const long long t_ready = p.dbg ? clock64() : 0;
        long long *Tbase = ((p.twin && (item & 1)) ? p.Tb : p.T) + ((int64_t)it.pos * WP_TILED + qd * 32) * KP_TILED + it.piece * Q_PW;
        for (int rd = r_begin; rd < r_end; ++rd) {
            const int i0 = rd * 8;
            const int n = Q_PW - i0 < 8 ? Q_PW - i0 : 8;  // 8, or 6 in the last round
            uint32_t r[Q_NV][8];
#pragma unroll
            for (int pl = 0; pl < Q_NV; ++pl)
#pragma unroll
                for (int j = 0; j < 8; ++j) r[pl][j] = 0u;
            if (!empty_list) {
                // eight columns per load (the columns past a plane's 102 points belong to the next plane or are the
                // piece's two spare columns: read, not used)
#pragma unroll
                for (int pl = 0; pl < Q_NV; ++pl) tmem_ld_32x32b_x8(tlane + (uint32_t)(pl * Q_PW + i0), r[pl]);
                tmem_wait_ld();
            }
            long long out[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int32_t a = (int32_t)r[0][j] + ((int32_t)r[1][j] << 8);
                const int32_t c = (int32_t)r[2][j] + ((int32_t)r[3][j] << 8);
                out[j] = (long long)a + ((long long)c << 16) + ((long long)(int32_t)r[4][j] << 32);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(xw + 512u * j + (uint32_t)(((lane + 8 * j) & 31) << 4)),
                             "l"(out[2 * j]), "l"(out[2 * j + 1])
                             : "memory");
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int bw = rb + 8 * i;  // boot within the warp's 32
                long long v0, v1;
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1)
                             : "r"(xw + 512u * rc + (uint32_t)(((bw + 8 * rc) & 31) << 4)) : "memory");
                // streaming stores: T is read back once, by the soft-max kernel, long after it has left the L2 -- it
                // should not push table rows out (an evict-first hint on the table loads, on the other hand, cost
                // 4 %: the rows of small counts are shared by thousands of genes and live in the L2)
                if (qd * 32 + bw < p.n_boot && 2 * rc < n)
                    __stcs(reinterpret_cast<longlong2 *>(Tbase + (int64_t)bw * KP_TILED + i0 + 2 * rc), make_longlong2(v0, v1));
            }
            __syncwarp();
        }
        tc_fence_before_sync();
        mbar_arrive(&sm.acc_empty);
        if (p.dbg && ewarp == 0 && lane == 0) {
            atomicAdd(p.dbg, (unsigned long long)(clock64() - t_ready));
            atomicAdd(p.dbg + 1, 1ull);
        }
    }
}

template <int Q_PGROUPS, bool HINT, int NS>
__global__ void __launch_bounds__(q_threads(Q_PGROUPS), 1) contract_i8_kernel(const I8Params p) {
    static_assert(NS >= 2 && NS <= Q_NS, "ring depth");
    constexpr int Q_PRODUCER_WARPS = 4 * Q_PGROUPS;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // stage buffers on a 1024-byte boundary (the swizzle pattern is a function of the shared-memory address bits)
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t stage0 = (raw + 1023u) & ~1023u;
    const uint32_t xbuf = stage0 + (uint32_t)NS * Q_STAGE_BYTES;
    I8Smem &sm = *reinterpret_cast<I8Smem *>(smem_raw + (stage0 - raw) + (size_t)NS * Q_STAGE_BYTES + Q_XBUF_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&sm.full[s], 4 * 32);                 // one cp.async completion arrival per thread of a producer group
            mbar_init(&sm.empty[s], 1);                     // one tcgen05.commit
        }
        mbar_init(&sm.acc_full, 1);
        mbar_init(&sm.acc_empty, Q_EPILOGUE_WARPS * 32);
        sm.abort = 0;
        mbar_fence_init();
    }
    if (warp == Q_PRODUCER_WARPS + Q_EPILOGUE_WARPS) tmem_alloc(&sm.tmem_base, Q_TMEM_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = sm.tmem_base;
    const int n_items = (p.n_pos * p.n_pieces) << p.twin;

    if (warp < Q_PRODUCER_WARPS)
        run_producer<Q_PGROUPS, HINT, NS>(p, sm, stage0, n_items, warp, lane);
    else if (warp < Q_PRODUCER_WARPS + Q_EPILOGUE_WARPS)
        run_epilogue(p, sm, xbuf, tmem, n_items, warp - Q_PRODUCER_WARPS, lane);
    else if (lane == 0)
        run_mma<NS>(p, sm, stage0, tmem, n_items);

    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (sm.abort && p.err && threadIdx.x == 0) atomicExch(p.err, 2);
    if (warp == Q_PRODUCER_WARPS + Q_EPILOGUE_WARPS) tmem_dealloc(tmem, Q_TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// Sentinel range of every (gene, boot): [max klo, min khi] over the list entries the boot drew.  One warp per gene, lane l
// owns boots 4 l .. 4 l + 3 (one 32-bit word of the entry's W row); klo / khi are kept as packed 16-bit pairs and
// combined with __vmaxu2 / __vminu2 under a mask made of the non-zero multiplicities.  Entries whose row has no sentinel
// (the common case) are skipped after one uniform load.  flag |= 4 when a drawn row is irregular or an intersection is
// empty.
__global__ void __launch_bounds__(256) sentinel_range_kernel(const int32_t *__restrict__ lst_row, const int32_t *__restrict__ lst_cell,
                                                             const int32_t *__restrict__ lst_len, const int32_t *__restrict__ order,
                                                             int64_t ld_lst, const uint32_t *__restrict__ row_range,
                                                             const int8_t *__restrict__ W8, int n_pos, int K, int n_boot,
                                                             uint32_t *__restrict__ SR, int32_t *__restrict__ flag) {
    const int pos = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pos >= n_pos) return;
    const int lane = threadIdx.x & 31;
    const int64_t gene = order ? order[pos] : pos;
    const int len = lst_len[gene];
    const int32_t *lr = lst_row + gene * ld_lst, *lc = lst_cell + gene * ld_lst;
    const uint32_t full = (uint32_t)(K - 1) << 16;
    uint32_t lo01 = 0u, lo23 = 0u, hi01 = 0xFFFFFFFFu, hi23 = 0xFFFFFFFFu;
    const uint32_t v01 = (4 * lane < n_boot ? 0xFFFFu : 0u) | (4 * lane + 1 < n_boot ? 0xFFFF0000u : 0u);
    const uint32_t v23 = (4 * lane + 2 < n_boot ? 0xFFFFu : 0u) | (4 * lane + 3 < n_boot ? 0xFFFF0000u : 0u);
    bool irregular = false;
    // the (row -> range) lookups of the next 32 entries are in flight while the current ones are combined
    auto fetch = [&](int e0, uint32_t &rr, int32_t &cell) {
        const int e = e0 + lane;
        rr = full;
        cell = 0;
        if (e < len) {
            rr = __ldg(row_range + lr[e]);
            cell = lc[e] & ~LIST_HOT_BIT;
        }
    };
    uint32_t rr_n = full;
    int32_t cell_n = 0;
    if (len > 0) fetch(0, rr_n, cell_n);
    for (int e0 = 0; e0 < len; e0 += 32) {
        const uint32_t rr = rr_n;
        const int32_t cell = cell_n;
        if (e0 + 32 < len) fetch(e0 + 32, rr_n, cell_n);
        // An entry whose klo is not above the smallest lower bound any boot holds so far, and whose khi is not below the
        // largest upper bound, cannot change any boot's range whatever its multiplicities are: it is skipped without
        // touching its W row.  The bounds tighten within the first few dozen entries (every boot draws 63 % of the
        // cells), after which almost every entry is skipped.  (Boots beyond n_boot hold no bounds: masked out.)
        const uint32_t lmin2 = __vminu2(lo01 | ~v01, lo23 | ~v23), hmax2 = __vmaxu2(hi01 & v01, hi23 & v23);
        const uint32_t gmin = __reduce_min_sync(0xffffffffu, min(lmin2 & 0xFFFFu, lmin2 >> 16));
        const uint32_t gmax = __reduce_max_sync(0xffffffffu, max(hmax2 & 0xFFFFu, hmax2 >> 16));
        unsigned need = __ballot_sync(0xffffffffu, rr != full && (rr == Q_RANGE_IRREGULAR || (rr & 0xFFFFu) > gmin ||
                                                                  (rr >> 16) < gmax));
        while (need) {  // four entries per round: their W words are loaded before any is used (the loop is latency-bound)
            uint32_t r[4], w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                r[u] = full;
                w[u] = 0u;
                if (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    r[u] = __shfl_sync(0xffffffffu, rr, src);
                    const int32_t c = __shfl_sync(0xffffffffu, cell, src);
                    w[u] = __ldg(reinterpret_cast<const uint32_t *>(W8 + (int64_t)c * Q_WB + 4 * lane));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (r[u] == Q_RANGE_IRREGULAR) {
                    irregular = true;
                    continue;
                }
                const uint32_t m = __vcmpne4(w[u], 0u);  // 0xFF per boot that drew this cell
                const uint32_t m01 = __byte_perm(m, 0u, 0x1100), m23 = __byte_perm(m, 0u, 0x3322);
                const uint32_t klo2 = (r[u] & 0xFFFFu) * 0x10001u, khi2 = (r[u] >> 16) * 0x10001u;
                lo01 = __vmaxu2(lo01, klo2 & m01);
                lo23 = __vmaxu2(lo23, klo2 & m23);
                hi01 = __vminu2(hi01, khi2 | ~m01);
                hi23 = __vminu2(hi23, khi2 | ~m23);
            }
        }
    }
    uint32_t out[4];
    bool bad = irregular;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t lo = ((j < 2 ? lo01 : lo23) >> (16 * (j & 1))) & 0xFFFFu;
        uint32_t hi = ((j < 2 ? hi01 : hi23) >> (16 * (j & 1))) & 0xFFFFu;
        if (hi > (uint32_t)(K - 1)) hi = (uint32_t)(K - 1);
        if (4 * lane + j < n_boot && lo > hi) bad = true;
        // points beyond khi: only real grid points matter, padding columns are never read by the soft-max
        out[j] = lo | (hi == (uint32_t)(K - 1) ? 0xFFFF0000u : hi << 16);
    }
    *reinterpret_cast<uint4 *>(SR + (int64_t)pos * Q_WB + 4 * lane) = make_uint4(out[0], out[1], out[2], out[3]);
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flag, 4);
}

// ------------------------------------------------------------------------------------------------
// softmax_i8_warp_kernel + softmax_i8_reduce_kernel: jp[gene, k] (+)= sum_b softmax_k(T[b, :])[k] / scale from the integer
// T tiles of contract_i8_kernel (src/jpmatLogBoot.cpp:264-269).  T[b, k] = 2^-29 * integer + Z[b, k], "log 0" (excluded
// from the soft-max) outside the (gene, boot) sentinel range.  A CTA owns one group of 13 boots; the eight groups' partial
// sums are added in group order by the reduce kernel: the result is deterministic.
constexpr int SP_GROUPS = 8, SP_ROWS = WP_TILED / SP_GROUPS;
static_assert(SP_GROUPS * SP_ROWS == WP_TILED && SP_ROWS * 32 == KP_TILED, "13 warps x 32 lanes = the 416 grid slots");
// softmax_i8_warp_kernel: a warp per gene.  A CTA owns one group of 13 boots, whose Z rows it keeps in shared memory
// (43 KB); each of its warps walks genes on its own and, for a gene, the 13 boots in ascending order: T row (the next
// boot's row is already in flight), soft-max pieces by shuffles (exp_nonpos, skipped warp-wide where a 32-point stretch
// lies > 746 nats below the row maximum: a joint posterior is sharply peaked), and the normalised row is added to 13
// accumulators per lane in ascending boot order.
constexpr int SW_WARPS = 8;
__global__ void __launch_bounds__(SW_WARPS * 32, 2)
softmax_i8_warp_kernel(const long long *__restrict__ T, const double *__restrict__ Z, const uint32_t *__restrict__ SR, int K,
                       int n_boot_pass, double scale, double *__restrict__ part, int n_pos) {
    constexpr int NJ = KP_TILED / 32;
    extern __shared__ double s_z[];  // [SP_ROWS][KP_TILED]
    const int group = blockIdx.x % SP_GROUPS, batch = blockIdx.x / SP_GROUPS, n_batch = gridDim.x / SP_GROUPS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b0 = group * SP_ROWS;
    const int nb = min(SP_ROWS, n_boot_pass - b0);  // boots of this group (<= 0: none)
    for (int i = threadIdx.x; i < SP_ROWS * KP_TILED; i += SW_WARPS * 32)
        s_z[i] = (Z && i / KP_TILED < nb) ? Z[(int64_t)b0 * KP_TILED + i] : 0.0;
    __syncthreads();
    const double q = 1.0 / (double)(1ll << Q_FRAC);
    for (int pos = batch * SW_WARPS + warp; pos < n_pos; pos += n_batch * SW_WARPS) {
        double acc[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[j] = 0.0;
        const long long *rows = T + ((int64_t)pos * WP_TILED + b0) * KP_TILED + lane;
        // the group's 13 range words, one per lane
        const uint32_t sr_l = (SR && lane < SP_ROWS) ? SR[(int64_t)pos * Q_WB + b0 + lane] : 0xFFFF0000u;
        long long nxt[NJ];
        if (nb > 0) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) nxt[j] = __ldcs(rows + 32 * j);
        }
        for (int bi = 0; bi < nb; ++bi) {
            long long cur[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) cur[j] = nxt[j];
            if (bi + 1 < nb) {
                const long long *r2 = rows + (int64_t)(bi + 1) * KP_TILED;
#pragma unroll
                for (int j = 0; j < NJ; ++j) nxt[j] = __ldcs(r2 + 32 * j);
            }
            const uint32_t sr = __shfl_sync(0xffffffffu, sr_l, bi);
            const int klo = (int)(sr & 0xFFFFu), span = min((int)(sr >> 16), K - 1) - klo;  // admissible: klo .. klo + span
            const double *zr = s_z + bi * KP_TILED + lane;
            double v[NJ], m = -INFINITY;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int k = lane + 32 * j;
                // some drawn row is "log 0" outside the range: the reference's T is a multiple of the sentinel there
                const double t = fma((double)cur[j], q, zr[32 * j]);
                v[j] = (unsigned)(k - klo) <= (unsigned)span && span >= 0 ? t : -INFINITY;
                m = v[j] > m ? v[j] : m;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double mo = __shfl_xor_sync(0xffffffffu, m, o);
                m = mo > m ? mo : m;
            }
            double sum = 0.0;
            uint32_t amask = 0u;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const double d = v[j] - m;
                v[j] = 0.0;
                if (__any_sync(0xffffffffu, d > -746.0)) {
                    v[j] = d > -INFINITY ? exp_nonpos(d) : 0.0;
                    amask |= 1u << j;
                }
                sum += v[j];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const double inv = 1.0 / (sum * scale);
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                if (amask & (1u << j)) acc[j] = __dadd_rn(acc[j], __dmul_rn(v[j], inv));
        }
        double *dst = part + ((int64_t)group * n_pos + pos) * KP_TILED + lane;
#pragma unroll
        for (int j = 0; j < NJ; ++j) dst[32 * j] = acc[j];
    }
}

__global__ void softmax_i8_reduce_kernel(const double *__restrict__ part, const int32_t *__restrict__ order, int K, int n_pos,
                                         double *__restrict__ jp, int64_t ld_jp, int accumulate) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_pos * KP_TILED) return;
    const int pos = (int)(idx / KP_TILED), k = (int)(idx - (int64_t)pos * KP_TILED);
    if (k >= K) return;
    double r = 0.0;
#pragma unroll
    for (int g = 0; g < SP_GROUPS; ++g) r += part[((int64_t)g * n_pos + pos) * KP_TILED + k];
    const int64_t gene = order ? order[pos] : pos;
    double *out = jp + gene * ld_jp + k;
    *out = accumulate ? *out + r : r;
}

// the integer T tiles as the FP64 values the reference would hold (tests: scde_b200_probe_contract_i8), in place
__global__ void finalize_t_kernel(long long *__restrict__ T, const uint32_t *__restrict__ SR, double sentinel, int64_t n) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int k = (int)(idx % KP_TILED);
    const int64_t row = idx / KP_TILED;
    const int b = (int)(row % WP_TILED);
    const int64_t pos = row / WP_TILED;
    const uint32_t sr = SR ? SR[pos * Q_WB + b] : 0xFFFF0000u;
    double t = (double)T[idx] * (1.0 / (double)(1ll << Q_FRAC));
    if (k < (int)(sr & 0xFFFFu) || k > (int)(sr >> 16)) t = sentinel;
    reinterpret_cast<double *>(T)[idx] = t;
}

// ------------------------------------------------------------------------------------------------
// fixed-point planes and non-sentinel range of every row of an FP64 table (the general row kernel writes FP64 only; the
// constant-theta fast kernel emits both itself, lp_table.cu): one warp per row
__global__ void __launch_bounds__(256) quantize_rows_kernel(const double *__restrict__ table, int ld_table, int K, int64_t n_rows,
                                                            int8_t *__restrict__ qtable, int64_t ldq,
                                                            uint32_t *__restrict__ row_range) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const double *src = table + row * ld_table;
    int8_t *dst = qtable + row * ldq;
    for (int64_t j = lane; j < ldq / 4; j += 32) reinterpret_cast<uint32_t *>(dst)[j] = 0u;
    __syncwarp();
    int n_ok = 0, kmin = 0x7fffffff, kmax = -1;
    for (int k = lane; k < K; k += 32) {
        const double v = src[k];
        if (!(v > -1.0e290)) continue;  // the "log 0" sentinel (or a difference to it): digits stay zero
        ++n_ok;
        kmin = min(kmin, k);
        kmax = max(kmax, k);
        long long x = __double2ll_rn(fmin(fmax(v, -1000.0), 1000.0) * (double)(1ll << Q_FRAC));
        const int pc = k / Q_PW, i = k - pc * Q_PW;
#pragma unroll
        for (int pl = 0; pl < Q_NV; ++pl) {
            const int d = (int)(int8_t)(x & 0xFF);
            dst[pc * Q_PIECE + pl * Q_PW + i] = (int8_t)d;
            x = (x - d) >> 8;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_ok += __shfl_xor_sync(0xffffffffu, n_ok, o);
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if (lane == 0) {
        uint32_t rr;
        if (n_ok == 0)
            rr = Q_RANGE_IRREGULAR;  // a row of sentinels only: let the FP64 kernel deal with it
        else if (kmax - kmin + 1 != n_ok)
            rr = Q_RANGE_IRREGULAR;
        else
            rr = (uint32_t)kmin | ((uint32_t)kmax << 16);
        row_range[row] = rr;
    }
}

// W (FP64 multiplicities, pass-major [pass][n_w_rows][108]) -> int8 rows of 128 bytes; flag |= 1 when a count > 127
__global__ void w_to_i8_kernel(const double *__restrict__ W, int64_t n_rows_total, int8_t *__restrict__ W8,
                               int32_t *__restrict__ flag) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows_total * Q_WB) return;
    const int64_t row = idx / Q_WB;
    const int col = (int)(idx - row * Q_WB);
    int v = 0;
    if (col < WP_TILED) {
        const double w = W[row * WS_TILED + col];
        if (w > 127.0) {
            atomicOr(flag, 1);
            v = 127;
        } else {
            v = (int)w;
        }
    }
    W8[idx] = (int8_t)v;
}

}  // namespace

cudaError_t launch_quantize_rows(const double *table, int ld_table, int K, int64_t n_rows, int8_t *qtable,
                                 uint32_t *row_range, cudaStream_t st) {
    if (n_rows <= 0) return cudaSuccess;
    if (K > ld_table || K > Q_MAX_K) return cudaErrorInvalidValue;
    quantize_rows_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, st>>>(table, ld_table, K, n_rows, qtable, q_row_bytes(K),
                                                                        row_range);
    return cudaGetLastError();
}

cudaError_t launch_w_to_i8(const double *W, int n_w_rows, int n_boot, int8_t *W8, int32_t *flag, cudaStream_t st) {
    const int passes = (n_boot + WP_TILED - 1) / WP_TILED;
    const int64_t rows = (int64_t)(passes > 0 ? passes : 1) * n_w_rows;
    const int64_t n = rows * Q_WB;
    w_to_i8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, rows, W8, flag);
    return cudaGetLastError();
}

bool contract_i8_supported(int K, int ld_table, int ld_lst, int n_draws) {
    return K >= 1 && K <= Q_MAX_K && ld_table == KP_TILED && (ld_lst % Q_ES) == 0 && n_draws <= 65000;
}

size_t contract_i8_range_words(int n_genes) { return (size_t)(n_genes > 0 ? n_genes : 1) * Q_WB; }

static I8Params make_params(const ContractI8Args &a, int g0, int n_pos, int pass, double *t_scratch) {
    I8Params p;
    p.qtable = a.qtable;
    p.ldq = a.ldq;
    p.lst_row = a.lists.row;
    p.lst_cell = a.lists.cell;
    p.lst_len = a.lists.len;
    p.order = a.lists.order ? a.lists.order + g0 : nullptr;
    p.ld_lst = a.lists.ld;
    p.W8 = a.W8 + (size_t)pass * a.n_w_rows * Q_WB;
    p.T = reinterpret_cast<long long *>(t_scratch);
    p.n_boot = (a.n_boot - pass * WP_TILED) < WP_TILED ? (a.n_boot - pass * WP_TILED) : WP_TILED;
    p.n_pos = n_pos;
    p.n_pieces = q_pieces(a.K);
    p.err = a.err;
    p.dbg = a.dbg;
    p.W8b = a.W8_twin ? a.W8_twin + (size_t)pass * a.n_w_rows * Q_WB : nullptr;
    p.Tb = reinterpret_cast<long long *>(a.t_twin);
    p.twin = (a.W8_twin && a.t_twin) ? 1 : 0;
    p.piece_major = a.item_order == 1;
    p.cold_evict_first = a.cold_evict_first;
    return p;
}

cudaError_t launch_sentinel_ranges(const ContractI8Args &a, int g0, int n_pos, int pass, uint32_t *sr_scratch,
                                   cudaStream_t st) {
    if (n_pos <= 0 || !a.row_range) return cudaSuccess;
    if (!sr_scratch) return cudaErrorInvalidValue;
    const I8Params p = make_params(a, g0, n_pos, pass, nullptr);
    const int nb = (a.n_boot - pass * WP_TILED) < WP_TILED ? (a.n_boot - pass * WP_TILED) : WP_TILED;
    sentinel_range_kernel<<<(unsigned)((n_pos + 7) / 8), 256, 0, st>>>(p.lst_row, p.lst_cell, p.lst_len, p.order, p.ld_lst,
                                                                       a.row_range, p.W8, n_pos, a.K, nb, sr_scratch, a.err);
    return cudaGetLastError();
}

cudaError_t launch_contract_i8_pass(const ContractI8Args &a, int g0, int n_pos, int pass, int n_sm, double *t_scratch,
                                    cudaStream_t st) {
    if (n_pos <= 0) return cudaSuccess;
    // ONE producer group (four warps, one per k-group of a stage).  Round 1 used two groups that took alternate stages:
    // no faster (74.37 against 74.34 ms per step at config 4, profiles/r02_experiments.txt) and unsafe -- a group that has
    // no stage of its own in a run of short items (lists of <= 32 entries) does not wait on anything, runs arbitrarily far
    // ahead of the MMA thread, and its next parity wait on a slot's "empty" barrier can then be satisfied by a phase two
    // uses back; with the piece-major item order (short items at the end of every phase, long ones at the start of the
    // next) that deadlocked scde.posteriors over 40 cells (tests: test_config2_..., test_short_and_long_lists_mixed...).
    // With one group every producer warp touches every stage, so the ring itself bounds its lead.
    constexpr int PG = 1;
    const I8Params p = make_params(a, g0, n_pos, pass, t_scratch);
    const int n_items = (n_pos * p.n_pieces) << p.twin;
    int grid = n_sm < n_items ? n_sm : n_items;
    if (p.twin) grid &= ~1;  // an even stride keeps every CTA on one joint and the pairs (2 i, 2 i + 1) together
    // function attributes are per device: set on every launch (a process may hold contexts on several GPUs)
    auto launch = [&](auto kernel, int ns) -> cudaError_t {
        const size_t smem = q_smem_bytes(ns);
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<grid, q_threads(PG), smem, st>>>(p);
        return cudaGetLastError();
    };
    if (a.hot_rank >= 0) return launch(contract_i8_kernel<PG, true, Q_NS>, Q_NS);
    // the shallow ring leaves 69 KB of shared memory and a third of the register file to a co-resident kernel
    if (a.ring_stages == 7) return launch(contract_i8_kernel<PG, false, 7>, 7);
    if (a.ring_stages == 8) return launch(contract_i8_kernel<PG, false, 8>, 8);
    return launch(contract_i8_kernel<PG, false, Q_NS>, Q_NS);
}

size_t softmax_i8_scratch_doubles(int n_pos) { return (size_t)SP_GROUPS * (n_pos > 0 ? n_pos : 1) * KP_TILED; }

cudaError_t launch_softmax_i8(const ContractI8Args &a, int g0, int n_pos, int pass, const double *t_scratch, const uint32_t *sr,
                              double *part_scratch, int n_sm, cudaStream_t st) {
    if (n_pos <= 0) return cudaSuccess;
    if (a.row_range && !sr) return cudaErrorInvalidValue;
    const int nb = (a.n_boot - pass * WP_TILED) < WP_TILED ? (a.n_boot - pass * WP_TILED) : WP_TILED;
    const double *Z = a.Z ? a.Z + (size_t)pass * WP_TILED * KP_TILED : nullptr;
    const size_t smem = sizeof(double) * SP_ROWS * KP_TILED;
    cudaError_t e = cudaFuncSetAttribute(softmax_i8_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int wb = n_sm * 2 / SP_GROUPS;  // two CTAs of eight warps per SM
    if (wb * SW_WARPS > n_pos) wb = (n_pos + SW_WARPS - 1) / SW_WARPS;
    if (wb < 1) wb = 1;
    softmax_i8_warp_kernel<<<wb * SP_GROUPS, SW_WARPS * 32, smem, st>>>(reinterpret_cast<const long long *>(t_scratch), Z,
                                                                        a.row_range ? sr : nullptr, a.K, nb, a.scale,
                                                                        part_scratch, n_pos);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int64_t n = (int64_t)n_pos * KP_TILED;
    softmax_i8_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part_scratch, a.lists.order ? a.lists.order + g0 : nullptr,
                                                                          a.K, n_pos, a.jp, a.ld_jp, pass > 0);
    return cudaGetLastError();
}

cudaError_t launch_finalize_t(const ContractI8Args &a, int n_pos, double *t_scratch, const uint32_t *sr, cudaStream_t st) {
    if (n_pos <= 0) return cudaSuccess;
    const int64_t n = (int64_t)n_pos * WP_TILED * KP_TILED;
    finalize_t_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<long long *>(t_scratch),
                                                                   a.row_range ? sr : nullptr, a.sentinel, n);
    return cudaGetLastError();
}

}  // namespace scde
