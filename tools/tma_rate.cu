// TMA engine throughput for small boxes of doubles (per SM, all 148 SMs active): cycles per tensor store / load as a function
// of the box shape {W nodes x H instances}.  One thread per CTA issues K copies back to back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_rate tools/tma_rate.cu && tools/tma_rate
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: stores, 1: loads.  Each CTA walks its own instance range; x advances by W (wraps at N)
__global__ void k(const __grid_constant__ CUtensorMap tm, int mode, int W, int H, int N, int K, int nthr, long long* cyc) {
    extern __shared__ __align__(128) double box[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
    }
    __syncthreads();
    // nthr issuing threads, one per warp
    if ((threadIdx.x & 31) != 0 || (threadIdx.x >> 5) >= nthr) return;
    const int w = threadIdx.x >> 5;
    const int y0 = blockIdx.x * H;
    long long t0 = clock64();
    if (mode == 0) {
        int x = w * W;
        for (int i = 0; i < K; i++) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(&tm), "r"(x % N), "r"(y0), "r"(s32(box)) : "memory");
            asm volatile("cp.async.bulk.commit_group;\n");
            x += nthr * W;
        }
        asm volatile("cp.async.bulk.wait_group 0;\n");
    } else {
        if (w == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(&bar)), "r"(K * W * H * 8));
            int x = 0;
            for (int i = 0; i < K; i++) {
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(s32(box)), "l"(&tm), "r"(x % N), "r"(y0), "r"(s32(&bar)) : "memory");
                x += W;
            }
            uint32_t done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }\n" : "=r"(done) : "r"(s32(&bar)) : "memory");
        }
    }
    long long t1 = clock64();
    if (w == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    const int N = 1296, H = 56, G = 148;
    const long long B = (long long)G * H;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    PFN_enc enc = (PFN_enc)sym;
    double* d;
    long long* cyc;
    cudaMalloc(&d, B * N * 8);
    cudaMemset(d, 0, B * N * 8);
    cudaMalloc(&cyc, G * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int K = 256;
    for (int mode = 0; mode < 2; mode++)
        for (int W : {4, 8, 16, 20, 36, 72, 144})
            for (int nthr : {1, 2, 4}) {
                if (mode == 1 && nthr > 1) continue;
                if (W * H * 8 > 100 * 1024) continue;
                CUtensorMap tm;
                cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)B}, strides[1] = {(cuuint64_t)N * 8};
                cuuint32_t bx[2] = {(cuuint32_t)W, (cuuint32_t)H}, es[2] = {1, 1};
                int r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r) { printf("encode failed %d\n", r); continue; }
                for (int rep = 0; rep < 2; rep++) k<<<G, 128, W * H * 8>>>(tm, mode, W, H, N, K, nthr, cyc);
                cudaError_t e = cudaDeviceSynchronize();
                std::vector<long long> h(G);
                cudaMemcpy(h.data(), cyc, G * 8, cudaMemcpyDeviceToHost);
                double avg = 0;
                for (auto v : h) avg += v;
                avg /= G;
                const double per = avg / ((double)K * (mode ? 1 : nthr)), rows = per / H, bpc = (double)W * H * 8 / per;
                printf("%s box {%3d x %d} issuers %d: %7.1f cycles/copy  %5.2f cycles/box-row  %6.1f B/cycle/SM  (%s)\n", mode ? "load " : "store", W, H, nthr, per, rows, bpc, cudaGetErrorString(e));
            }
    return 0;
}
