// HBM throughput of the block-sparse kernel's ACCESS PATTERN without any arithmetic: every CTA streams tiles of TI instances
// through a ring of TMA boxes {W nodes x TI instances} (loads, one producer thread) and writes them back as boxes
// {WO nodes x TI} (TMA tensor stores, one storing thread), batch [B][N] row-major as in config 3 (N = 1296, B = 65536).
// Answers: is ~4.3 TB/s the ceiling of 160-byte box rows / 32-byte store rows at a 10 KB pitch, or of the kernel structure?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_pattern tools/tma_pattern.cu && tools/tma_pattern
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint64_t* bar, uint32_t par) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }\n"
                     : "=r"(done) : "r"(s32(bar)), "r"(par) : "memory");
}
constexpr int MAXR = 32;

// mode bit0: loads, bit1: stores
__global__ void k(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, int mode, int W, int WO,
                  int TI, int N, long long B, int R, int PF) {
    extern __shared__ __align__(128) double ring[];
    __shared__ uint64_t full[MAXR], empty[MAXR];
    const int boxd = W * TI;
    if (threadIdx.x == 0) {
        for (int s = 0; s < R; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n");
    }
    __syncthreads();
    const long long ntiles = B / TI;
    const int nbt = (N + W - 1) / W;
    if (threadIdx.x == 0) {  // loader
        int slot = 0, round = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
            for (int j = 0; j < nbt; j++) {
                if (PF > 0 && (mode & 1)) {  // L2 prefetch of the box PF positions ahead in this CTA's stream (HBM -> L2, no smem)
                    int jp = j + PF;
                    long long tp = tile;
                    while (jp >= nbt) jp -= nbt, tp += gridDim.x;
                    if (tp < ntiles)
                        asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(&tin), "r"(jp * W), "r"((int)(tp * TI)) : "memory");
                }
                if (round > 0) mwait(&empty[slot], (round - 1) & 1);
                if (mode & 1) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(&full[slot])), "r"(boxd * 8) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
                                     "r"(s32(ring + (size_t)slot * boxd)), "l"(&tin), "r"(j * W), "r"((int)(tile * TI)), "r"(s32(&full[slot])) : "memory");
                } else {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s32(&full[slot])) : "memory");
                }
                if (++slot == R) slot = 0, round++;
            }
    } else if (threadIdx.x == 32) {  // storer: box j of the tile leaves as W / WO store boxes; slots are handed back two boxes later
        int slot = 0, round = 0, pend = 0, rel = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
            for (int j = 0; j < nbt; j++) {
                mwait(&full[slot], round & 1);
                if (mode & 2) {
                    for (int x = 0; x < W; x += WO)
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(&tout),
                                     "r"(j * W + x), "r"((int)(tile * TI)), "r"(s32(ring + (size_t)slot * boxd + (size_t)x * TI)) : "memory");
                    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                    if (++pend > 2) {
                        asm volatile("cp.async.bulk.wait_group.read 2;\n" ::: "memory");
                        pend--;
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s32(&empty[rel])) : "memory");
                        if (++rel == R) rel = 0;
                    }
                } else {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s32(&empty[slot])) : "memory");
                }
                if (++slot == R) slot = 0, round++;
            }
        asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    }
}

int main() {
    const int N = 1296, G = 148;
    const long long B = 65536;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    PFN_enc enc = (PFN_enc)sym;
    double *din, *dout;
    cudaMalloc(&din, B * N * 8);
    cudaMalloc(&dout, B * N * 8);
    cudaMemset(din, 0, B * N * 8);
    cudaMemset(dout, 0, B * N * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    auto mk = [&](CUtensorMap* tm, double* base, int w, int h) {
        cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)B}, strides[1] = {(cuuint64_t)N * 8};
        cuuint32_t bx[2] = {(cuuint32_t)w, (cuuint32_t)h}, es[2] = {1, 1};
        return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    struct Cfg { int W, WO, TI, KB, PF; };
    // ring budget in KB; the block-sparse kernel has ~45 KB of boxes in flight at {20 x 56}
    std::vector<Cfg> cfgs = {{20, 4, 56, 45, 0}, {20, 4, 56, 45, 4}, {20, 4, 56, 45, 8}, {20, 4, 56, 45, 16}, {20, 4, 56, 45, 32},
                             {20, 4, 56, 27, 0}, {20, 4, 56, 27, 8}, {20, 4, 56, 27, 16}, {20, 4, 56, 90, 0}, {20, 4, 56, 90, 8},
                             {20, 4, 56, 90, 16}, {20, 20, 56, 90, 0}, {20, 20, 56, 180, 0}, {20, 4, 40, 32, 0}, {20, 4, 40, 32, 12},
                             {20, 4, 40, 64, 0}, {20, 4, 40, 64, 12}, {144, 144, 16, 180, 0}};
    for (int mode : {3, 1})
        for (const Cfg& c : cfgs) {
            CUtensorMap tin, tout;
            if (mk(&tin, din, c.W, c.TI) || mk(&tout, dout, c.WO, c.TI)) { printf("encode failed\n"); continue; }
            const int boxb = c.W * c.TI * 8;
            const int R = std::min(MAXR, std::max(3, c.KB * 1024 / boxb));
            const size_t smem = (size_t)R * boxb;
            float best = 1e30f;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                k<<<G, 64, smem>>>(tin, tout, mode, c.W, c.WO, c.TI, N, B, R, c.PF);
                cudaEventRecord(e1);
                if (cudaEventSynchronize(e1) != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); break; }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                best = std::min(best, ms);
            }
            const double bytes = (double)B * N * 8 * ((mode & 1) + ((mode >> 1) & 1));
            printf("%s load {%3d x %3d} store {%3d x %3d} ring %2d boxes (%3zu KB) L2 prefetch %2d boxes ahead: %.4f ms  %7.1f GB/s\n",
                   mode == 3 ? "copy " : mode == 1 ? "read " : "write", c.W, c.TI, c.WO, c.TI, R, smem / 1024, c.PF, best, bytes / best / 1e6);
        }
    return 0;
}
