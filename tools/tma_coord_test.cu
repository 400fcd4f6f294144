// Does a TMA tile-mode box of doubles need a 16-byte aligned origin (even innermost coordinate)?  Loads and stores boxes of
// {20 x 8} doubles at even and odd innermost coordinates and checks the data.   nvcc -arch=sm_100a tools/tma_coord_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, int x, int y, int xo,
                  double* dbg) {
    __shared__ __align__(128) double box[8 * 20];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(&bar)), "r"(8 * 20 * 8));
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(s32(box)),
                     "l"(&tin), "r"(x), "r"(y), "r"(s32(&bar))
                     : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }\n" : "=r"(done) : "r"(s32(&bar)) : "memory");
        for (int i = 0; i < 160; i++) dbg[i] = box[i];
        asm volatile("fence.proxy.async.shared::cta;\n");
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(&tout), "r"(xo), "r"(y), "r"(s32(box)) : "memory");
        asm volatile("cp.async.bulk.commit_group;\n");
        asm volatile("cp.async.bulk.wait_group 0;\n");
    }
}

int main(int argc, char** argv) {
    const int ax = argc > 2 ? atoi(argv[1]) : -1, axo = argc > 2 ? atoi(argv[2]) : -1;
    const int C = 202, R = 64;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    PFN_enc enc = (PFN_enc)sym;
    std::vector<double> h(R * C);
    for (int i = 0; i < R * C; i++) h[i] = i;
    double *din, *dout, *dbg;
    cudaMalloc(&din, R * C * 8); cudaMalloc(&dout, R * C * 8); cudaMalloc(&dbg, 160 * 8);
    cudaMemcpy(din, h.data(), R * C * 8, cudaMemcpyHostToDevice);
    CUtensorMap tin, tout;
    cuuint64_t dims[2] = {C, R}, strides[1] = {C * 8};
    cuuint32_t bx[2] = {20, 8}, es[2] = {1, 1};
    int r1 = enc(&tin, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, din, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int r2 = enc(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, dout, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d %d\n", r1, r2);
    int fails = 0;
    for (int x : {ax}) {
        for (int xo : {axo}) {
            if (xo < 0) continue;
            cudaMemset(dout, 0, R * C * 8);
            k<<<1, 32>>>(tin, tout, x, 16, xo, dbg);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<double> b(160), o(R * C);
            cudaMemcpy(b.data(), dbg, 160 * 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(o.data(), dout, R * C * 8, cudaMemcpyDeviceToHost);
            int badl = 0, bads = 0;
            for (int r = 0; r < 8; r++)
                for (int c = 0; c < 20; c++) {
                    double want = (x + c < C) ? (double)((16 + r) * C + x + c) : 0.0;
                    if (b[r * 20 + c] != want) badl++;
                    if (xo + c < C && o[(16 + r) * C + xo + c] != b[r * 20 + c]) bads++;
                }
            long nz = 0;
            for (int i = 0; i < R * C; i++) nz += o[i] != 0.0;
            printf("load x=%d store x=%d: err=%s load_mismatch=%d store_mismatch=%d nonzero_out=%ld\n", x, xo, cudaGetErrorString(e), badl, bads, nz);
            fails += badl + bads + (e != cudaSuccess);
        }
    }
    printf(fails ? "FAIL\n" : "PASS: arbitrary innermost coordinates are fine\n");
    return 0;
}
