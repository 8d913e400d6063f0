// gather_bench.cu -- how fast can a B200 gather random contiguous pieces of table rows from HBM?
// The tcgen05 contraction (csrc/contract_i8.cu) reads, per list entry, one contiguous piece of a random table row;
// this microbenchmark measures the rate of exactly that access pattern as a function of the piece size, the row
// stride and whether neighbouring CTAs read neighbouring pieces of the same rows at about the same time.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/gather_bench.cu -o gpurun_out/gather_bench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t mix(uint32_t a, uint32_t b) {
    uint64_t x = ((uint64_t)a << 32) | b;
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return (uint32_t)x;
}

// every CTA walks `n_pos` random rows; a warp takes positions warp, warp + n_warps, ...; per position it reads
// `granule` bytes at row * stride + chunk * granule with 16-byte loads, lanes side by side
template <int UNROLL>
__global__ void __launch_bounds__(1024, 2) gather_kernel(const uint4 *__restrict__ buf, uint64_t n_rows, int stride16,
                                                         int granule16, int n_chunk, int share, int n_pos,
                                                         uint32_t *sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const uint32_t seq = share ? blockIdx.x / n_chunk : blockIdx.x;
    const int chunk_fixed = share ? blockIdx.x % n_chunk : -1;
    uint32_t acc = 0;
    for (int p0 = warp * UNROLL; p0 < n_pos; p0 += n_warps * UNROLL) {
        uint4 v[UNROLL][4];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t h = mix(seq, (uint32_t)(p0 + u));
            const uint64_t row = (uint64_t)h % n_rows;
            const int chunk = chunk_fixed >= 0 ? chunk_fixed : (int)(mix(h, 7u) % (uint32_t)n_chunk);
            const uint4 *src = buf + row * (uint64_t)stride16 + (uint64_t)chunk * granule16;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int o = lane + 32 * j;
                v[u][j] = make_uint4(0, 0, 0, 0);
                if (o < granule16) v[u][j] = __ldg(src + o);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc ^= v[u][j].x ^ v[u][j].y ^ v[u][j].z ^ v[u][j].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

// sibling-prefetch variant: the CTA demand-loads its own piece (512 B) of each row and at the same time asks the L2 to
// fetch the other piece(s) of the same 1 KB (or 2 KB) block, which a neighbouring CTA walking the same rows will demand
// a little later (mode 1: prefetch.global.L2 per 128-byte line, 2: cp.async.bulk.prefetch.L2 of the whole sibling
// piece by one lane, 3: prefetch.global.L2 per 32-byte sector)
__global__ void __launch_bounds__(1024, 2) sibling_kernel(const uint4 *__restrict__ buf, uint64_t n_rows, int stride16,
                                                           int n_chunk, int block_pieces, int mode, int n_pos,
                                                           uint32_t *sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const uint32_t seq = blockIdx.x / n_chunk;
    const int chunk = blockIdx.x % n_chunk;
    uint32_t acc = 0;
    for (int p0 = warp * 2; p0 < n_pos; p0 += n_warps * 2) {
        uint4 v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t h = mix(seq, (uint32_t)(p0 + u));
            const uint64_t row = (uint64_t)h % n_rows;
            const uint4 *rowp = buf + row * (uint64_t)stride16;
            v[u] = __ldg(rowp + chunk * 32 + lane);
            const int first = chunk / block_pieces * block_pieces;  // pieces of the same aligned block
            for (int sib = first; sib < first + block_pieces; ++sib) {
                if (sib == chunk) continue;
                const uint4 *sp = rowp + sib * 32;
                if (mode == 1) {
                    if (lane < 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + lane * 8));
                } else if (mode == 2) {
                    if (lane == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(sp), "r"(512));
                } else if (mode == 3) {
                    if (lane < 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + lane * 2));
                }
            }
        }
        acc ^= v[0].x ^ v[0].y ^ v[0].z ^ v[0].w ^ v[1].x ^ v[1].y ^ v[1].z ^ v[1].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

// lockstep variant: clusters of `csize` CTAs walk the same random rows, CTA r of the cluster copying piece r (512 B) of
// each row with cp.async into shared memory; with sync = 1 a cluster barrier after every group of copies keeps the CTAs
// within a few hundred cycles of each other, so the pieces of one aligned block are requested at (almost) the same time
// by different SMs.  dup = 1: all CTAs of the cluster copy the SAME piece (second and later touches can hit the L2).
__global__ void __launch_bounds__(1024, 1) lockstep_kernel(const uint4 *__restrict__ buf, uint64_t n_rows, int stride16,
                                                            int csize, int sync, int dup, int n_iter) {
    extern __shared__ uint4 sbuf[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t seq = blockIdx.x / csize;
    const int piece = dup ? 0 : blockIdx.x % csize;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sbuf);
    for (int it = 0; it < n_iter; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t h = mix(seq, (uint32_t)((it * 32 + warp) * 4 + u));
            const uint64_t row = (uint64_t)h % n_rows;
            const uint4 *src = buf + row * (uint64_t)stride16 + piece * 32 + lane;
            const uint32_t dst = s0 + (uint32_t)((((it & 1) * 4 + u) * 1024 + threadIdx.x) * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        if (sync) {
            asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__global__ void stream_kernel(const uint4 *__restrict__ buf, uint64_t n16, uint32_t *sink) {
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(buf + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    const uint64_t bytes = 24ull << 30;
    uint4 *buf;
    uint32_t *sink;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) return 1;
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto time_it = [&](auto launch) {
        launch();
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int r = 0; r < 3; ++r) launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        return ms / 3;
    };
    {
        const float ms = time_it([&] { stream_kernel<<<148 * 16, 512>>>(buf, bytes / 16, sink); });
        printf("{\"case\": \"stream read\", \"GBps\": %.1f}\n", bytes / ms / 1e6);
    }
    struct Case {
        int stride, granule, n_chunk, share;
    };
    const Case cases[] = {
        {2496, 480, 5, 1}, {2496, 480, 5, 0}, {2048, 512, 4, 1}, {2048, 512, 4, 0}, {2048, 1024, 2, 1}, {2048, 1024, 2, 0},
        {2048, 2048, 1, 0}, {2048, 256, 8, 0}, {2048, 128, 16, 0}, {4096, 4096 / 2, 2, 0}, {2560, 512, 5, 1}, {2560, 512, 5, 0},
    };
    for (const Case &c : cases) {
        const uint64_t n_rows = bytes / c.stride;
        const int grid = 148 * 2 / c.n_chunk * c.n_chunk;
        const int n_pos = (int)((40ull << 30) / c.granule / grid);  // ~40 GB per launch
        for (int unroll = 2; unroll <= 4; unroll += 2) {
            float ms;
            if (unroll == 2)
                ms = time_it([&] { gather_kernel<2><<<grid, 1024>>>(buf, n_rows, c.stride / 16, c.granule / 16, c.n_chunk, c.share, n_pos, sink); });
            else
                ms = time_it([&] { gather_kernel<4><<<grid, 1024>>>(buf, n_rows, c.stride / 16, c.granule / 16, c.n_chunk, c.share, n_pos, sink); });
            printf("{\"case\": \"gather\", \"stride\": %d, \"granule\": %d, \"pieces\": %d, \"neighbours_share_rows\": %d, \"unroll\": %d, "
                   "\"GBps\": %.1f}\n",
                   c.stride, c.granule, c.n_chunk, c.share, unroll, (double)n_pos * grid * c.granule / ms / 1e6);
        }
    }
    struct SCase {
        int stride, n_chunk, block_pieces;
    };
    const SCase scases[] = {{2048, 4, 2}, {2048, 4, 4}, {2048, 2, 2}};
    for (const SCase &c : scases)
        for (int mode = 0; mode <= 3; ++mode) {
            const uint64_t n_rows = bytes / c.stride;
            const int grid = 148 * 2 / c.n_chunk * c.n_chunk;
            const int n_pos = (int)((40ull << 30) / 512 / grid);
            const float ms = time_it([&] { sibling_kernel<<<grid, 1024>>>(buf, n_rows, c.stride / 16, c.n_chunk, c.block_pieces, mode, n_pos, sink); });
            printf("{\"case\": \"sibling prefetch\", \"stride\": %d, \"pieces_walked\": %d, \"block_pieces\": %d, \"mode\": %d, \"demand_GBps\": %.1f}\n",
                   c.stride, c.n_chunk, c.block_pieces, mode, (double)n_pos * grid * 512 / ms / 1e6);
        }
    struct LCase {
        int stride, csize, sync, dup;
    };
    const LCase lcases[] = {{2048, 2, 0, 0}, {2048, 2, 1, 0}, {2048, 4, 0, 0}, {2048, 4, 1, 0}, {2048, 1, 0, 0},
                            {2048, 2, 0, 1}, {2048, 2, 1, 1}};
    cudaFuncSetAttribute(lockstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    for (const LCase &c : lcases) {
        const uint64_t n_rows = bytes / c.stride;
        const int grid = 148 / c.csize * c.csize;
        const int n_iter = (int)((30ull << 30) / (64 * 1024) / grid);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(1024);
        cfg.dynamicSmemBytes = 128 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = c.csize;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const float ms = time_it([&] {
            cudaLaunchKernelEx(&cfg, lockstep_kernel, (const uint4 *)buf, n_rows, c.stride / 16, c.csize, c.sync, c.dup, n_iter);
        });
        printf("{\"case\": \"lockstep cp.async 512 B pieces\", \"stride\": %d, \"cluster\": %d, \"sync\": %d, \"dup\": %d, \"demand_GBps\": %.1f}\n",
               c.stride, c.csize, c.sync, c.dup, (double)n_iter * grid * 64 * 1024 / ms / 1e6);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("cuda error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
