// contract_i8.cu -- the bootstrap joint posterior (src/jpmatLogBoot.cpp:216-275, :460-499) on the 5th-generation tensor
// cores: tcgen05.mma.kind::i8 with 32-bit integer accumulators in tensor memory.
//
// Per gene the reference computes  T[b, k] = sum_c W[c, b] * lp[c, x[g, c], k]  (W = how often boot b drew cell c) and
// then jp[g, :] = (1/B) sum_b softmax_k T[b, :].  boot_contract.cu evaluates that sum on the FP64 pipe (DMMA), where it
// is bound by the 37 TFLOP/s FP64 rate.  Here the table is held in FIXED POINT instead: in the zero-base form every
// table value D = lp(x) - lp(0) (or lp(x) itself) lies in [-751, 751] -- log of a normalised double -- or is the
// reference's "log 0" sentinel (-DBL_MAX/n/1.1, :127,204).  D is rounded to a multiple of 2^-29 and split into five
// signed radix-256 digits (int8 planes 0..4, value = 2^-29 * sum_p 256^p d_p); a sixth plane holds the sentinel
// indicator.  W is a small non-negative integer (<= 127 is checked).  Every product and every sum is then EXACT in
// int32 (|sum| <= 128 * draws), the planes are recombined in int64 and converted to FP64 once per (boot, grid point),
// so the only error is the table rounding: |T error| <= draws * 2^-30 in the worst case and about
// sqrt(2 * draws) * 2^-29 / sqrt(12) typically (5e-8 at 5000 cells per group) -- inside the 1e-6 tolerance on
// log-posteriors, and independent of the summation order, so the result is deterministic by construction.
//
// Kernel: persistent, one CTA per SM, three warp roles, no CTA-wide barrier in the steady state.
//   * work item = (gene, grid chunk of up to 80 points).  One item accumulates all six planes of its chunk:
//     D[boot (128 lanes, 104 real)][plane * w + i] in 6 * 80 = 480 of the SM's 512 tensor-memory columns.
//   * producers (4 warps): gather the list entries' table pieces (480 contiguous bytes per entry: the table is stored
//     [row][chunk][plane][w]) and W rows (128 bytes) with 16-byte cp.async straight into the UMMA "interleave" (no
//     swizzle) canonical layout for MN-major operands -- lane (kk = lane & 7, piece) writes 16 bytes of entry kk to
//     [k-group][piece][kk][16 B], so one warp instruction fills 512 contiguous bytes of shared memory (no bank
//     conflicts) from eight 64-byte runs of global memory.  Completion is signalled by cp.async.mbarrier.arrive; a
//     10-stage ring of 32 entries (19 KB per stage) keeps ~190 KB in flight per SM.
//   * MMA (1 thread): per stage two tcgen05.mma (M = 128 boots, N = 240 = three planes, K = 32 entries), then
//     tcgen05.commit onto the stage's "empty" barrier; after the last stage a commit onto the accumulator barrier.
//   * epilogue (4 warps, one per 32-lane quarter of tensor memory): tcgen05.ld the six planes of 8 grid points at a
//     time, recombine in int64, add the zero-count base Z[b, k] (FP64), store T[b, k]; softmax_avg_kernel
//     (boot_contract.cu) finishes the gene.  Producers keep prefetching the next item's stages meanwhile.
// Roofline: HBM/L2 gather bandwidth (608 bytes per visited (gene, cell) pair and chunk); the tensor pipe needs
// 2 x 120 cycles per 32 entries, about a quarter of the time the gather takes.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include <cfloat>
#include <cmath>

namespace scde {
namespace {

using namespace ptx;

constexpr int Q_ES = 32;                         // list entries per stage = K of one tcgen05.mma.kind::i8
constexpr int Q_NS = 10;                         // ring depth
constexpr int Q_A_BYTES = Q_ES * Q_WB;           // W tile of a stage: 4096
constexpr int Q_B_BYTES = Q_ES * 512;            // table tile of a stage: 480 bytes per entry, padded to 4 x 128
constexpr int Q_STAGE_BYTES = Q_A_BYTES + Q_B_BYTES;  // 20480, a multiple of the 1024-byte swizzle atom
constexpr int Q_PGROUPS = 2;                     // producer warp groups; group g gathers the stages s = g (mod Q_PGROUPS) of an item
constexpr int Q_PRODUCER_WARPS = 4 * Q_PGROUPS, Q_EPILOGUE_WARPS = 4;
constexpr int Q_THREADS = (Q_PRODUCER_WARPS + Q_EPILOGUE_WARPS + 1) * 32;  // + the MMA warp
constexpr int Q_TMEM_COLS = 512;
constexpr long long Q_WATCHDOG_CYCLES = 4000000000ll;  // a barrier wait longer than ~2 s aborts the kernel (err = 2)
constexpr int Q_LAYOUT_SW128 = 0, Q_LAYOUT_INTERLEAVE = 1;

struct I8Smem {
    uint64_t full[Q_NS];
    uint64_t empty[Q_NS];
    uint64_t acc_full, acc_empty;
    uint32_t tmem_base;
    volatile int abort;
};
constexpr size_t Q_SMEM_BYTES = 1024 /* alignment slack */ + (size_t)Q_NS * Q_STAGE_BYTES + sizeof(I8Smem);

struct I8Params {
    const int8_t *qtable;  // [rows][ldq]: row = [chunk][plane][w_chunk]
    int64_t ldq;
    const int32_t *lst_row, *lst_cell, *lst_len, *order;
    int64_t ld_lst;
    const int8_t *W8;  // this pass: [n_w_rows][128]
    const double *Z;   // this pass: [104][416] or NULL
    double *T;         // [n_pos][104][416]
    double sentinel;   // added once per drawn "log 0" entry, as the FP64 path does
    int n_pos;         // genes in this launch: positions [0, n_pos) of `order`
    int n_chunks;      // grid chunks per gene; all but the last are Q_CW wide
    int w_last;        // width of the last chunk (multiple of 16, <= Q_CW)
    int32_t *err;      // device flag: 2 = watchdog abort
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

struct Item {
    int pos, chunk, w;
};
__device__ __forceinline__ Item decode_item(const I8Params &p, int item) {
    // all full-width chunks first (genes heaviest first), the narrow last chunks at the end: they even out the tail
    const int nfull = p.n_chunks - 1, n_first = p.n_pos * nfull;
    Item it;
    if (item < n_first) {
        it.pos = item / nfull;
        it.chunk = item - it.pos * nfull;
        it.w = Q_CW;
    } else {
        it.pos = item - n_first;
        it.chunk = nfull;
        it.w = p.w_last;
    }
    return it;
}

// ---- producers ------------------------------------------------------------------------------------------------------
// Warp w gathers entries 8w .. 8w+7 (one k-group of the MMA) of every stage.  List entries are read one coalesced load
// per four stages and three groups ahead (a load issued at stage 4t is first used at stage 4t + 8), so the gather never
// waits on the list: with a one-stage look-ahead the kernel ran at the latency of that dependent load (long-scoreboard
// stalls were 66 % of the producers' cycles, profiles/r01p).
//
// SW128 layout (canonical 128-byte-swizzle layout of an MN-major operand): the bytes of one entry are split into runs of
// 128 (8 pieces of 16 bytes), a run of entry kk sits at [run][k-group][kk][128 B] and its piece j at 16 * (j ^ kk).
// Lane (r = lane >> 3, j = lane & 7) copies piece j of entries r and r + 4: a warp instruction moves four 128-byte runs of
// global memory into four 128-byte lines of shared memory -- coalesced on both sides, no bank conflicts.
// INTERLEAVE layout (no swizzle): piece j of entry kk at [k-group][j][kk][16 B]; lane (kk = lane & 7, q = lane >> 3) copies
// pieces q, q + 4, ...  Verified first; kept as a cross-check (its shared-memory side serialises: the four pieces of one
// 64-byte global run land 128 bytes apart, i.e. in the same banks).
template <int LAYOUT>
__device__ __forceinline__ void run_producer(const I8Params &p, I8Smem &sm, uint32_t stage0, int n_items, int warp, int lane) {
    // A single warp issues the ~180 instructions of a stage (address arithmetic, ten cp.async, barrier traffic) at well
    // under one per cycle, so four warps alone ran the gather at half the rate the memory system sustains
    // (profiles/r01s): Q_PGROUPS groups of four warps take alternate stages.
    const int grp = warp >> 2, kg = warp & 3;  // k-group of the MMA this warp fills: entries 8 kg .. 8 kg + 7 of a stage
    int64_t q0 = 0;                            // stages of the items before this one
    const int l7 = lane & 7, l3 = lane >> 3;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        const int64_t gene = p.order ? p.order[it.pos] : it.pos;
        const int nst = (p.lst_len[gene] + Q_ES - 1) / Q_ES;
        const int npieces = Q_NP * (it.w >> 4);
        const int8_t *qbase = p.qtable + (int64_t)it.chunk * (Q_NP * Q_CW);
        const int64_t lbase = gene * p.ld_lst + kg * 8 + l7;
        // this warp's stages are s = grp + Q_PGROUPS * i; lane (l3, l7) holds entry l7 of own-stage 4 t + l3 for the
        // register groups t, t + 1, t + 2
        int32_t er[3] = {0, 0, 0}, ec[3] = {0, 0, 0};
        auto load_group = [&](int t, int32_t &row, int32_t &cell) {
            const int s = grp + Q_PGROUPS * (4 * t + l3);
            row = 0;
            cell = 0;
            if (s < nst) {
                row = p.lst_row[lbase + (int64_t)s * Q_ES];
                cell = p.lst_cell[lbase + (int64_t)s * Q_ES];
            }
        };
        load_group(0, er[0], ec[0]);
        load_group(1, er[1], ec[1]);
        load_group(2, er[2], ec[2]);
        int i = 0;
        for (int s = grp; s < nst; s += Q_PGROUPS, ++i) {
            if (i > 0 && (i & 3) == 0) {
                er[0] = er[1];
                ec[0] = ec[1];
                er[1] = er[2];
                ec[1] = ec[2];
                load_group((i >> 2) + 2, er[2], ec[2]);
            }
            const int64_t q = q0 + s;
            const int slot = (int)(q % Q_NS);
            const uint32_t fill = (uint32_t)(q / Q_NS);
            const uint32_t sA = stage0 + (uint32_t)slot * Q_STAGE_BYTES;
            const uint32_t sB = sA + Q_A_BYTES;
            const int g4 = (i & 3) << 3;
            if (LAYOUT == Q_LAYOUT_SW128) {
                const int32_t row0 = __shfl_sync(0xffffffffu, er[0], g4 | l3), row1 = __shfl_sync(0xffffffffu, er[0], g4 | (l3 + 4));
                const int32_t cel0 = __shfl_sync(0xffffffffu, ec[0], g4 | l3), cel1 = __shfl_sync(0xffffffffu, ec[0], g4 | (l3 + 4));
                if (fill > 0 && !wait_or_abort(sm, &sm.empty[slot], (fill - 1) & 1u)) return;
                const int8_t *src0 = qbase + (int64_t)row0 * p.ldq + l7 * 16;
                const int8_t *src1 = qbase + (int64_t)row1 * p.ldq + l7 * 16;
                // entry kk = l3 (and l3 + 4): line kk of this warp's k-group, piece l7 at 16 * (l7 ^ kk)
                const uint32_t d0 = (uint32_t)kg * 1024u + (uint32_t)l3 * 128u + (uint32_t)((l7 ^ l3) << 4);
                const uint32_t d1 = (uint32_t)kg * 1024u + (uint32_t)(l3 + 4) * 128u + (uint32_t)((l7 ^ (l3 + 4)) << 4);
                for (int run = 0; run * 8 + l7 < npieces; ++run) {
                    cp_async16(sB + (uint32_t)run * 4096u + d0, src0 + run * 128);
                    cp_async16(sB + (uint32_t)run * 4096u + d1, src1 + run * 128);
                }
                cp_async16(sA + d0, p.W8 + (int64_t)cel0 * Q_WB + l7 * 16);
                cp_async16(sA + d1, p.W8 + (int64_t)cel1 * Q_WB + l7 * 16);
            } else {
                const int32_t row = __shfl_sync(0xffffffffu, er[0], g4 | l7);
                const int32_t cell = __shfl_sync(0xffffffffu, ec[0], g4 | l7);
                if (fill > 0 && !wait_or_abort(sm, &sm.empty[slot], (fill - 1) & 1u)) return;
                const int8_t *src = qbase + (int64_t)row * p.ldq;
                const uint32_t dB = sB + (uint32_t)kg * (uint32_t)(npieces * 128) + (uint32_t)l7 * 16u;
                for (int j = l3; j < npieces; j += 4) cp_async16(dB + (uint32_t)j * 128u, src + j * 16);
                const int8_t *wsrc = p.W8 + (int64_t)cell * Q_WB;
                const uint32_t dA = sA + (uint32_t)kg * 1024u + (uint32_t)l7 * 16u;
                cp_async16(dA + (uint32_t)l3 * 128u, wsrc + l3 * 16);
                cp_async16(dA + (uint32_t)(l3 + 4) * 128u, wsrc + (l3 + 4) * 16);
            }
            cp_async_mbar_arrive_noinc(&sm.full[slot]);
        }
        q0 += nst;
    }
}

// ---- MMA issuer (one thread) ------------------------------------------------------------------------------------------
template <int LAYOUT>
__device__ __forceinline__ void run_mma(const I8Params &p, I8Smem &sm, uint32_t stage0, uint32_t tmem, int n_items) {
    int64_t q = 0;
    uint32_t n_done = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
        const Item it = decode_item(p, item);
        const int64_t gene = p.order ? p.order[it.pos] : it.pos;
        const int nst = (p.lst_len[gene] + Q_ES - 1) / Q_ES;
        const int ntot = Q_NP * it.w;  // accumulator columns of this item (480 for a full chunk)
        // N <= 256 per instruction.  SW128: the second MMA has to start on a 128-byte run of the entry -> 256 + 224;
        // INTERLEAVE: any multiple of 16 -> two halves
        int n1, n2;
        if (LAYOUT == Q_LAYOUT_SW128) {
            n1 = ntot > 256 ? 256 : ntot;
            n2 = ntot - n1;
        } else {
            n1 = ntot > 256 ? ntot / 2 : ntot;
            n2 = ntot - n1;
        }
        const uint32_t idesc1 = umma_idesc_s8_mn(128, n1), idesc2 = umma_idesc_s8_mn(128, n2 > 0 ? n2 : 16);
        const uint32_t stride_k_b = (uint32_t)(Q_NP * (it.w >> 4)) * 128u;  // INTERLEAVE: bytes between 8-entry groups of B
        if (n_done > 0) {  // the epilogue has drained the previous item's accumulators
            if (!wait_or_abort(sm, &sm.acc_empty, (n_done - 1) & 1u)) return;
            tc_fence_after_sync();
        }
        for (int s = 0; s < nst; ++s, ++q) {
            const int slot = (int)(q % Q_NS);
            if (!wait_or_abort(sm, &sm.full[slot], (uint32_t)(q / Q_NS) & 1u)) return;
            fence_proxy_async_smem();
            tc_fence_after_sync();
            const uint32_t sA = stage0 + (uint32_t)slot * Q_STAGE_BYTES;
            const uint32_t sB = sA + Q_A_BYTES;
            if (LAYOUT == Q_LAYOUT_SW128) {
                const uint64_t da = umma_desc_sw128(sA, 4096u, 1024u);
                umma_s8(tmem, da, umma_desc_sw128(sB, 4096u, 1024u), idesc1, s > 0);
                if (n2 > 0) umma_s8(tmem + 256u, da, umma_desc_sw128(sB + 2u * 4096u, 4096u, 1024u), idesc2, s > 0);
            } else {
                const uint64_t da = umma_desc_nosw(sA, 128u, 1024u, false);
                umma_s8(tmem, da, umma_desc_nosw(sB, 128u, stride_k_b, false), idesc1, s > 0);
                if (n2 > 0)
                    umma_s8(tmem + (uint32_t)n1, da, umma_desc_nosw(sB + (uint32_t)(n1 >> 4) * 128u, 128u, stride_k_b, false),
                            idesc2, s > 0);
            }
            umma_commit(&sm.empty[slot]);  // frees the slot once these MMAs have read it
        }
        if (nst > 0)
            umma_commit(&sm.acc_full);
        else
            mbar_arrive(&sm.acc_full);
    }
}

// ---- epilogue: warp (4 + qd) owns tensor-memory lanes 32 qd .. 32 qd + 31 ----------------------------------------------
__device__ __forceinline__ void run_epilogue(const I8Params &p, I8Smem &sm, uint32_t tmem, int n_items, int warp, int lane) {
    const int qd = warp & 3;
    const int b = qd * 32 + lane;  // boot of this thread
    const uint32_t tlane = tmem + ((uint32_t)(qd * 32) << 16);
    const double scale = 1.0 / (double)(1ll << Q_FRAC);
    uint32_t n_done = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
        const Item it = decode_item(p, item);
        const int64_t gene = p.order ? p.order[it.pos] : it.pos;
        const bool empty_list = p.lst_len[gene] <= 0;
        while (!mbar_try_wait(&sm.acc_full, n_done & 1u)) {  // an item takes tens of microseconds: sleep, do not spin
            __nanosleep(500);
            if (sm.abort) return;
        }
        tc_fence_after_sync();
        double *Trow = p.T + ((int64_t)it.pos * WP_TILED + b) * KP_TILED + it.chunk * Q_CW;
        const double *Zrow = p.Z ? p.Z + (int64_t)b * KP_TILED + it.chunk * Q_CW : nullptr;
        for (int i0 = 0; i0 < it.w; i0 += 8) {
            uint32_t r[Q_NP][8];
            if (!empty_list) {
#pragma unroll
                for (int pl = 0; pl < Q_NP; ++pl) tmem_ld_32x32b_x8(tlane + (uint32_t)(pl * it.w + i0), r[pl]);
                tmem_wait_ld();
            } else {
#pragma unroll
                for (int pl = 0; pl < Q_NP; ++pl)
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[pl][j] = 0u;
            }
            if (b < WP_TILED) {
                double out[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    long long s = 0;
#pragma unroll
                    for (int pl = Q_NV - 1; pl >= 0; --pl) s = s * 256 + (long long)(int32_t)r[pl][j];
                    double t = (double)s * scale;
                    if (Zrow) t += Zrow[i0 + j];
                    const int32_t ns = (int32_t)r[Q_NV][j];  // draws that hit a "log 0" entry at this grid point
                    if (ns != 0) t = fma((double)ns, p.sentinel, t);
                    out[j] = t;
                }
#pragma unroll
                for (int j = 0; j < 8; j += 2)
                    *reinterpret_cast<double2 *>(Trow + i0 + j) = make_double2(out[j], out[j + 1]);
            }
        }
        tc_fence_before_sync();
        mbar_arrive(&sm.acc_empty);
    }
}

template <int LAYOUT>
__global__ void __launch_bounds__(Q_THREADS, 1) contract_i8_kernel(const I8Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // stage buffers on a 1024-byte boundary (the swizzle pattern is a function of the shared-memory address bits)
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t stage0 = (raw + 1023u) & ~1023u;
    I8Smem &sm = *reinterpret_cast<I8Smem *>(smem_raw + (stage0 - raw) + (size_t)Q_NS * Q_STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < Q_NS; ++s) {
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
    const int n_items = p.n_pos * p.n_chunks;

    if (warp < Q_PRODUCER_WARPS)
        run_producer<LAYOUT>(p, sm, stage0, n_items, warp, lane);
    else if (warp < Q_PRODUCER_WARPS + Q_EPILOGUE_WARPS)
        run_epilogue(p, sm, tmem, n_items, warp, lane);
    else if (lane == 0)
        run_mma<LAYOUT>(p, sm, stage0, tmem, n_items);

    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (sm.abort && p.err && threadIdx.x == 0) atomicExch(p.err, 2);
    if (warp == Q_PRODUCER_WARPS + Q_EPILOGUE_WARPS) tmem_dealloc(tmem, Q_TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// fixed-point planes of the table: one thread per four consecutive grid points of a row
__global__ void quantize_rows_kernel(const double *__restrict__ table, int ld_table, int kp, int64_t n_rows,
                                     int8_t *__restrict__ qtable, int64_t ldq) {
    const int tpr = kp >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * tpr) return;
    const int64_t row = idx / tpr;
    const int k = (int)(idx - row * tpr) * 4;
    const double *src = table + row * ld_table + k;
    const int c = k / Q_CW, i = k - c * Q_CW;
    const int w = min(Q_CW, kp - c * Q_CW);
    uint32_t word[Q_NP];
#pragma unroll
    for (int pl = 0; pl < Q_NP; ++pl) word[pl] = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double v = (k + j < ld_table) ? src[j] : 0.0;
        if (!(v > -1.0e290)) {  // the "log 0" sentinel (or a difference to it)
            word[Q_NV] |= 1u << (8 * j);
        } else {
            long long x = __double2ll_rn(fmin(fmax(v, -1000.0), 1000.0) * (double)(1ll << Q_FRAC));
#pragma unroll
            for (int pl = 0; pl < Q_NV; ++pl) {
                const int d = (int)(int8_t)(x & 0xFF);
                word[pl] |= (uint32_t)(d & 0xFF) << (8 * j);
                x = (x - d) >> 8;
            }
        }
    }
    int8_t *dst = qtable + row * ldq + (int64_t)c * (Q_NP * Q_CW) + i;
#pragma unroll
    for (int pl = 0; pl < Q_NP; ++pl) *reinterpret_cast<uint32_t *>(dst + pl * w) = word[pl];
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

int q_row_bytes(int K) { return Q_NP * round_up(K, 16); }

cudaError_t launch_quantize_rows(const double *table, int ld_table, int K, int64_t n_rows, int8_t *qtable,
                                 cudaStream_t st) {
    if (n_rows <= 0) return cudaSuccess;
    const int kp = round_up(K, 16);
    if (kp > ld_table) return cudaErrorInvalidValue;
    const int64_t n = n_rows * (kp >> 2);
    quantize_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(table, ld_table, kp, n_rows, qtable, q_row_bytes(K));
    return cudaGetLastError();
}

cudaError_t launch_w_to_i8(const double *W, int n_w_rows, int n_boot, int8_t *W8, int32_t *flag, cudaStream_t st) {
    const int passes = (n_boot + WP_TILED - 1) / WP_TILED;
    const int64_t rows = (int64_t)(passes > 0 ? passes : 1) * n_w_rows;
    const int64_t n = rows * Q_WB;
    w_to_i8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, rows, W8, flag);
    return cudaGetLastError();
}

bool contract_i8_supported(int K, int ld_table, int ld_lst) {
    return K >= 1 && K <= KP_TILED && ld_table == KP_TILED && (ld_lst % Q_ES) == 0;
}

cudaError_t launch_contract_i8_pass(const ContractI8Args &a, int g0, int n_pos, int pass, int n_sm, double *t_scratch,
                                    cudaStream_t st) {
    if (n_pos <= 0) return cudaSuccess;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(contract_i8_kernel<Q_LAYOUT_SW128>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Q_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(contract_i8_kernel<Q_LAYOUT_INTERLEAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)Q_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int kp = round_up(a.K, 16);
    const int n_chunks = (kp + Q_CW - 1) / Q_CW;
    I8Params p;
    p.qtable = a.qtable;
    p.ldq = a.ldq;
    p.lst_row = a.lists.row;
    p.lst_cell = a.lists.cell;
    p.lst_len = a.lists.len;
    p.order = a.lists.order ? a.lists.order + g0 : nullptr;
    p.ld_lst = a.lists.ld;
    p.W8 = a.W8 + (size_t)pass * a.n_w_rows * Q_WB;
    p.Z = a.Z ? a.Z + (size_t)pass * WP_TILED * KP_TILED : nullptr;
    p.T = t_scratch;
    p.sentinel = a.sentinel;
    p.n_pos = n_pos;
    p.n_chunks = n_chunks;
    p.w_last = kp - (n_chunks - 1) * Q_CW;
    p.err = a.err;
    const int n_items = n_pos * n_chunks;
    const int grid = n_sm < n_items ? n_sm : n_items;
    if (a.layout == Q_LAYOUT_INTERLEAVE)
        contract_i8_kernel<Q_LAYOUT_INTERLEAVE><<<grid, Q_THREADS, Q_SMEM_BYTES, st>>>(p);
    else
        contract_i8_kernel<Q_LAYOUT_SW128><<<grid, Q_THREADS, Q_SMEM_BYTES, st>>>(p);
    return cudaGetLastError();
}

}  // namespace scde
