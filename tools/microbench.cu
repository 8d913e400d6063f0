// microbench.cu -- FP64 roofs of the device the contraction kernel runs on: DFMA (vector pipe) and DMMA
// (mma.sync.m8n8k4.f64) throughput, and shared-memory broadcast behaviour of LDS.128.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int CHAINS>
__global__ void dfma_kernel(double *sink, int iters) {
    double a[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const double x = 1.0000000001, y = 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
}

template <int TILES>
__global__ void dmma_kernel(double *sink, int iters) {
    double c[TILES][2];
#pragma unroll
    for (int i = 0; i < TILES; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < TILES; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < TILES; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

// DMMA and DFMA issued from the same warp: do the FP64 tensor sub-pipe and the FP64 vector pipe run concurrently?
template <int TILES, int CHAINS>
__global__ void dmma_dfma_kernel(double *sink, int iters) {
    double c[TILES][2], f[CHAINS];
#pragma unroll
    for (int i = 0; i < TILES; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) f[i] = threadIdx.x * 1e-9 + i;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    const double x = 1.0000000001, y = 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < TILES; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
            for (int j = 0; j < CHAINS / TILES; ++j) f[i * (CHAINS / TILES) + j] = fma(f[i * (CHAINS / TILES) + j], x, y);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < TILES; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += f[i];
    if (s == 123.456) sink[0] = s;
}

// LDS.128 patterns: mode 0 = all lanes same address (broadcast), 1 = 8 distinct 16B chunks (lane>>2), 2 = 32 distinct
__global__ void lds_kernel(double *sink, int iters, int mode) {
    __shared__ double2 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_double2(i, -i);
    __syncthreads();
    int lane = threadIdx.x & 31;
    int idx = mode == 0 ? 0 : (mode == 1 ? (lane >> 2) : lane);
    double2 acc = make_double2(0, 0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double vx, vy;
            unsigned addr = (unsigned)__cvta_generic_to_shared(&buf[(idx + u * 32 + it) & 1023]);
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(addr));
            acc.x += vx;
            acc.y += vy;
        }
    }
    if (acc.x == 123.456) sink[0] = acc.y;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double *sink;
    CK(cudaMalloc(&sink, 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d", p.name, sms, p.clockRate / 1000);
    {
        const int iters = 20000, blocks = sms * 8, thr = 256;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            dfma_kernel<16><<<blocks, thr>>>(sink, iters);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf(", \"dfma_tflops\": %.3f", 2.0 * 16 * iters * (double)thr * blocks / (ms * 1e-3) / 1e12);
        // 12 warps/SM (the contraction kernel's occupancy), 52 chains
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            dfma_kernel<52><<<sms, 384>>>(sink, iters);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf(", \"dfma_tflops_12warps\": %.3f", 2.0 * 52 * iters * 384.0 * sms / (ms * 1e-3) / 1e12);
    }
    {
        const int iters = 20000, blocks = sms * 4, thr = 256;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            dmma_kernel<8><<<blocks, thr>>>(sink, iters);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
        }
        // per warp instruction: 8*8*4 FMAs = 512 flops
        printf(", \"dmma_m8n8k4_tflops\": %.3f", 512.0 * 8 * iters * (double)(thr / 32) * blocks / (ms * 1e-3) / 1e12);
    }
    {   // DMMA at the contraction kernel's occupancies: (warps per SM, independent tiles per warp)
        const int iters = 4000;
#define DMMA_CASE(TILES, THR)                                                                                     \
        for (int rep = 0; rep < 3; ++rep) {                                                                       \
            cudaEventRecord(e0);                                                                                  \
            dmma_kernel<TILES><<<sms, THR>>>(sink, iters);                                                        \
            cudaEventRecord(e1);                                                                                  \
            CK(cudaEventSynchronize(e1));                                                                         \
            cudaEventElapsedTime(&ms, e0, e1);                                                                    \
        }                                                                                                         \
        printf(", \"dmma_tflops_%dwarps_%dtiles\": %.3f", THR / 32, TILES,                                         \
               512.0 * TILES * iters * (double)(THR / 32) * sms / (ms * 1e-3) / 1e12);
        DMMA_CASE(26, 384)
        DMMA_CASE(33, 384)
        DMMA_CASE(21, 512)
        DMMA_CASE(13, 768)
        DMMA_CASE(42, 256)
        DMMA_CASE(26, 128)
        DMMA_CASE(8, 1024)
    }
    {   // both FP64 pipes at once, from the same warps: per loop trip TILES DMMA (512 flop each) + CHAINS DFMA (64 flop each)
        const int iters = 4000;
#define MIX_CASE(TILES, CHAINS)                                                                                    \
        for (int rep = 0; rep < 3; ++rep) {                                                                       \
            cudaEventRecord(e0);                                                                                  \
            dmma_dfma_kernel<TILES, CHAINS><<<sms, 384>>>(sink, iters);                                           \
            cudaEventRecord(e1);                                                                                  \
            CK(cudaEventSynchronize(e1));                                                                         \
            cudaEventElapsedTime(&ms, e0, e1);                                                                    \
        }                                                                                                         \
        printf(", \"mix_dmma%d_dfma%d_tflops\": %.3f", TILES, CHAINS,                                              \
               (512.0 * TILES + 64.0 * CHAINS) * iters * 12.0 * sms / (ms * 1e-3) / 1e12);
        MIX_CASE(16, 16)
        MIX_CASE(16, 32)
        MIX_CASE(16, 64)
        MIX_CASE(16, 128)
    }
    for (int mode = 0; mode < 3; ++mode) {
        const int iters = 20000, blocks = sms, thr = 512;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            lds_kernel<<<blocks, thr>>>(sink, iters, mode);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
        }
        // warp-level LDS.128 instructions per second per SM, in instructions per clock at the reported clock
        double inst = 8.0 * iters * (thr / 32);
        printf(", \"lds128_mode%d_inst_per_us_per_sm\": %.1f", mode, inst / (ms * 1e3));
    }
    printf("}\n");
    return 0;
}
