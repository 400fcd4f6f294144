// How fast can ONE warp per SM sub-partition feed DMMA.8x8x4 when the A operands come from 64-bit LDS (the block-sparse kernel's
// inner loop: per fragment 7 A loads + 1 B load + 7 DMMAs on 7 accumulators)?   cycles per DMMA for several variants.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_lds_rate tools/dmma_lds_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: operands in registers (no LDS); 1: A and B from LDS, loaded one fragment ahead (software pipelined);
// 2: like 1 plus a dependent index load (ring code -> address) two fragments ahead; 3: A from LDS right before use (no pipelining)
template <int MT, int MODE>
__global__ void k(double* out, long long* cyc, int rounds, int warps_active) {
    __shared__ double s[5120];
    __shared__ int idx[256];
    for (int i = threadIdx.x; i < 5120; i += blockDim.x) s[i] = 1e-3 * (i % 97);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) idx[i] = (i * 37) % 64;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= warps_active) return;
    const int g = lane >> 2, t = lane & 3;
    double acc[MT][2];
    for (int m = 0; m < MT; m++) acc[m][0] = acc[m][1] = 0.0;
    const double* su = s + g * 20 + t;
    double a_cur[MT], b_cur = s[lane];
    int o_nxt = idx[0] * 4;
    for (int m = 0; m < MT; m++) a_cur[m] = su[m * 160];
    long long t0 = clock64();
    for (int f = 1; f <= rounds; f++) {
        if (MODE == 0) {
            for (int m = 0; m < MT; m++) dmma(acc[m][0], acc[m][1], a_cur[m], b_cur);
        } else if (MODE == 3) {
            const int o = idx[f & 255] * 4;
            const double b = s[3072 + (f & 63) * 32 + lane];
            for (int m = 0; m < MT; m++) {
                const double a = su[m * 160 + o];
                dmma(acc[m][0], acc[m][1], a, b);
            }
        } else {
            const double b_nxt = s[3072 + (f & 63) * 32 + lane];
            double a_nxt[MT];
            const int o = MODE == 2 ? o_nxt : (f & 63) * 4;
#pragma unroll
            for (int m = 0; m < MT; m++) a_nxt[m] = su[m * 160 + o];
            if (MODE == 2) o_nxt = idx[(f + 1) & 255] * 4;
#pragma unroll
            for (int m = 0; m < MT; m++) dmma(acc[m][0], acc[m][1], a_cur[m], b_cur);
            b_cur = b_nxt;
#pragma unroll
            for (int m = 0; m < MT; m++) a_cur[m] = a_nxt[m];
        }
    }
    long long t1 = clock64();
    double r = 0;
    for (int m = 0; m < MT; m++) r += acc[m][0] + acc[m][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (lane == 0) cyc[blockIdx.x * 32 + warp] = t1 - t0;
}
template <int MT, int MODE>
void run(const char* name, int warps) {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 32 * 8);
    const int rounds = 4096;
    for (int r = 0; r < 2; r++) k<MT, MODE><<<148, 256, 0>>>(out, cyc, rounds, warps);
    cudaDeviceSynchronize();
    long long h[148 * 32];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int b = 0; b < 148; b++) avg += h[b * 32];
    avg /= 148;
    printf("%-50s MT=%d warps/SM=%d: %6.1f cycles per fragment round, %5.1f cycles per DMMA per warp, %5.2f cycles per DMMA per sub-partition\n", name, MT, warps, avg / rounds, avg / rounds / MT, avg / rounds / MT / ((warps + 3) / 4));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int warps : {1, 4, 8}) {
        run<7, 0>("registers only", warps);
        run<8, 0>("registers only", warps);
        run<7, 1>("LDS one fragment ahead", warps);
        run<7, 2>("LDS one ahead + dependent index two ahead", warps);
        run<7, 3>("LDS right before use", warps);
        run<4, 1>("LDS one fragment ahead", warps);
        run<14, 1>("LDS one fragment ahead", warps);
    }
    return 0;
}
