// microbench.cu -- measures the B200 denominators the roofline needs but MEASURED_PEAKS.json lacks:
//   * FP64 vector FMA peak (DFMA) and FP64 tensor-core peak for every mma.sync f64 shape (DMMA)
//   * fragment-layout self-check of each DMMA shape against a host product
//   * streaming copy bandwidth: LDG.128/STG.128 grid-stride copy and TMA bulk (UBLKCP) copy through smem
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include "../summationbypartsoperatorsextra.jl_b200/csrc/ptx.cuh"

using namespace sbp;

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

// ---------------- peaks -----------------------------------------------------------------
template <int ILP>
__global__ void k_dfma(double* out, int iters, double s) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-9 + i;
    double m = 1.0 + s;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], m, s);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r += acc[i];
    if (r == 12345.678) out[0] = r;
}

template <int SHAPE, int ILP>
__global__ void k_dmma(double* out, int iters, double s) {
    int lane = threadIdx.x & 31;
    double a8[8], b4[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a8[i] = 1e-3 * (lane + i) * s;
#pragma unroll
    for (int i = 0; i < 4; i++) b4[i] = 1e-3 * (lane - i) * s;
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) c[i][j] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (SHAPE == 884) dmma884(c[i][0], c[i][1], a8[0], b4[0]);
            if (SHAPE == 1684) {
                double a2[2] = {a8[0], a8[1]};
                dmma1684(c[i], a2, b4[0]);
            }
            if (SHAPE == 1688) {
                double a4[4] = {a8[0], a8[1], a8[2], a8[3]};
                double b2[2] = {b4[0], b4[1]};
                dmma1688(c[i], a4, b2);
            }
            if (SHAPE == 16816) dmma16816(c[i], a8, b4);
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) r += c[i][j];
    if (r == 12345.678) out[0] = r;
}

// ---------------- DMMA layout checks ------------------------------------------------------
// A: M x K row-major, Bm: K x 8 (element (k,n) at k*8+n), C: M x 8
template <int SHAPE>
__global__ void k_layout(const double* A, const double* Bm, double* C) {
    int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    if (SHAPE == 884) {
        double c0 = 0, c1 = 0;
        dmma884(c0, c1, A[g * 4 + t], Bm[t * 8 + g]);
        C[g * 8 + 2 * t] = c0;
        C[g * 8 + 2 * t + 1] = c1;
    } else {
        const int K = SHAPE == 1684 ? 4 : (SHAPE == 1688 ? 8 : 16);
        double c[4] = {0, 0, 0, 0};
        if (SHAPE == 1684) {
            double a[2] = {A[g * K + t], A[(g + 8) * K + t]};
            dmma1684(c, a, Bm[t * 8 + g]);
        } else if (SHAPE == 1688) {
            double a[4] = {A[g * K + t], A[(g + 8) * K + t], A[g * K + t + 4], A[(g + 8) * K + t + 4]};
            double b[2] = {Bm[t * 8 + g], Bm[(t + 4) * 8 + g]};
            dmma1688(c, a, b);
        } else {
            double a[8], b[4];
            for (int j = 0; j < 4; j++) {
                a[2 * j] = A[g * K + t + 4 * j];
                a[2 * j + 1] = A[(g + 8) * K + t + 4 * j];
                b[j] = Bm[(t + 4 * j) * 8 + g];
            }
            dmma16816(c, a, b);
        }
        C[g * 8 + 2 * t] = c[0];
        C[g * 8 + 2 * t + 1] = c[1];
        C[(g + 8) * 8 + 2 * t] = c[2];
        C[(g + 8) * 8 + 2 * t + 1] = c[3];
    }
}

template <int SHAPE>
double check_layout() {
    const int M = SHAPE == 884 ? 8 : 16;
    const int K = SHAPE == 884 ? 4 : (SHAPE == 1684 ? 4 : (SHAPE == 1688 ? 8 : 16));
    std::vector<double> A(M * K), Bm(K * 8), C(M * 8), R(M * 8, 0.0);
    for (int i = 0; i < M * K; i++) A[i] = sin(1.0 + i * 0.37);
    for (int i = 0; i < K * 8; i++) Bm[i] = cos(2.0 + i * 0.11);
    for (int m = 0; m < M; m++)
        for (int n = 0; n < 8; n++)
            for (int k = 0; k < K; k++) R[m * 8 + n] += A[m * K + k] * Bm[k * 8 + n];
    double *dA, *dB, *dC;
    CK(cudaMalloc(&dA, A.size() * 8));
    CK(cudaMalloc(&dB, Bm.size() * 8));
    CK(cudaMalloc(&dC, C.size() * 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bm.data(), Bm.size() * 8, cudaMemcpyHostToDevice));
    k_layout<SHAPE><<<1, 32>>>(dA, dB, dC);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));
    double err = 0;
    for (int i = 0; i < M * 8; i++) err = std::max(err, fabs(C[i] - R[i]));
    cudaFree(dA), cudaFree(dB), cudaFree(dC);
    return err;
}

// ---------------- copy bandwidth -----------------------------------------------------------
__global__ void k_copy_ldg(const double2* __restrict__ src, double2* __restrict__ dst, size_t n2) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 v = ldg_stream2(reinterpret_cast<const double*>(src + i));
        stg_stream2(reinterpret_cast<double*>(dst + i), v.x, v.y);
    }
}
__global__ void k_copy_ldg4(const double2* __restrict__ src, double2* __restrict__ dst, size_t n2) {
    // 4 independent 128-bit loads in flight per thread
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        double2 v0 = ldg_stream2((const double*)(src + i));
        double2 v1 = ldg_stream2((const double*)(src + i + stride));
        double2 v2 = ldg_stream2((const double*)(src + i + 2 * stride));
        double2 v3 = ldg_stream2((const double*)(src + i + 3 * stride));
        stg_stream2((double*)(dst + i), v0.x, v0.y);
        stg_stream2((double*)(dst + i + stride), v1.x, v1.y);
        stg_stream2((double*)(dst + i + 2 * stride), v2.x, v2.y);
        stg_stream2((double*)(dst + i + 3 * stride), v3.x, v3.y);
    }
    for (; i < n2; i += stride) {
        double2 v = ldg_stream2((const double*)(src + i));
        stg_stream2((double*)(dst + i), v.x, v.y);
    }
}

// TMA bulk copy through shared memory: STAGES buffers of TILE bytes per CTA, one producer thread.
template <int TILE, int STAGES>
__global__ void __launch_bounds__(128) k_copy_tma(const char* __restrict__ src, char* __restrict__ dst, size_t ntiles) {
    extern __shared__ __align__(128) char smem[];
    __shared__ uint64_t full[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        size_t first = blockIdx.x, stride = gridDim.x;
        // prologue
        int issued = 0;
        size_t t_load = first;
        for (; issued < STAGES && t_load < ntiles; issued++, t_load += stride) {
            mbar_arrive_expect_tx(&full[issued], TILE);
            bulk_g2s(smem + (size_t)issued * TILE, src + t_load * TILE, TILE, &full[issued]);
        }
        int s = 0;
        uint32_t ph = 0;
        for (size_t t = first; t < ntiles; t += stride) {
            mbar_wait(&full[s], ph);
            bulk_s2g(dst + t * TILE, smem + (size_t)s * TILE, TILE);
            bulk_commit();
            // before re-filling this stage the store must have finished READING smem
            bulk_wait_read<0>();
            if (t_load < ntiles) {
                mbar_arrive_expect_tx(&full[s], TILE);
                bulk_g2s(smem + (size_t)s * TILE, src + t_load * TILE, TILE, &full[s]);
                t_load += stride;
            }
            if (++s == STAGES) s = 0, ph ^= 1;
        }
        bulk_wait<0>();
    }
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    f();
    CK(cudaDeviceSynchronize());
    std::vector<float> ts;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* dout;
    CK(cudaMalloc(&dout, 64));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);

    // ---- layout checks
    printf(", \"layout_err\": {\"m8n8k4\": %.3e, \"m16n8k4\": %.3e, \"m16n8k8\": %.3e, \"m16n8k16\": %.3e}",
           check_layout<884>(), check_layout<1684>(), check_layout<1688>(), check_layout<16816>());

    // ---- DFMA peak
    {
        const int iters = 4096, ILP = 8, thr = 256, blocks = sms * 8;
        float ms = time_ms([&] { k_dfma<ILP><<<blocks, thr>>>(dout, iters, 1e-9); }, 7);
        double flops = 2.0 * iters * ILP * (double)thr * blocks;
        printf(", \"dfma_tflops\": %.3f", flops / ms / 1e9);
        // longer run (sustained, ~1 s)
        int long_iters = (int)(iters * (1000.0 / ms));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_dfma<ILP><<<blocks, thr>>>(dout, long_iters, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float lms;
        cudaEventElapsedTime(&lms, e0, e1);
        printf(", \"dfma_tflops_sustained_1s\": %.3f", 2.0 * long_iters * ILP * (double)thr * blocks / lms / 1e9);
    }
    // ---- DMMA peaks
    auto dmma_run = [&](int shape, int warps_per_block, int blocks_per_sm) -> double {
        const int iters = 2048, ILP = 8;
        int thr = warps_per_block * 32, blocks = sms * blocks_per_sm;
        float ms = 0;
        double fl = 0;
        switch (shape) {
            case 884:
                ms = time_ms([&] { k_dmma<884, ILP><<<blocks, thr>>>(dout, iters, 1e-9); }, 7);
                fl = 2.0 * 8 * 8 * 4;
                break;
            case 1684:
                ms = time_ms([&] { k_dmma<1684, ILP><<<blocks, thr>>>(dout, iters, 1e-9); }, 7);
                fl = 2.0 * 16 * 8 * 4;
                break;
            case 1688:
                ms = time_ms([&] { k_dmma<1688, ILP><<<blocks, thr>>>(dout, iters, 1e-9); }, 7);
                fl = 2.0 * 16 * 8 * 8;
                break;
            case 16816:
                ms = time_ms([&] { k_dmma<16816, ILP><<<blocks, thr>>>(dout, iters, 1e-9); }, 7);
                fl = 2.0 * 16 * 8 * 16;
                break;
        }
        return fl * iters * ILP * (double)warps_per_block * blocks / ms / 1e9;
    };
    printf(", \"dmma_tflops\": {");
    int shapes[4] = {884, 1684, 1688, 16816};
    for (int si = 0; si < 4; si++) {
        printf("%s\"%d\": {", si ? ", " : "", shapes[si]);
        int cfgs[4][2] = {{4, 1}, {8, 1}, {8, 2}, {16, 2}};
        for (int c = 0; c < 4; c++)
            printf("%s\"w%d_b%d\": %.3f", c ? ", " : "", cfgs[c][0], cfgs[c][1], dmma_run(shapes[si], cfgs[c][0], cfgs[c][1]));
        printf("}");
    }
    printf("}");

    // ---- copy bandwidth (1 GiB each way -> 2 GiB traffic), buffers >> L2
    {
        size_t bytes = (size_t)1 << 30;
        char *a, *b;
        CK(cudaMalloc(&a, bytes));
        CK(cudaMalloc(&b, bytes));
        CK(cudaMemset(a, 1, bytes));
        CK(cudaMemset(b, 2, bytes));
        size_t n2 = bytes / 16;
        printf(", \"copy_gbs\": {");
        float ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); }, 9);
        printf("\"cudaMemcpyD2D\": %.1f", 2.0 * bytes / ms / 1e6);
        for (int mult : {4, 8, 16, 32}) {
            ms = time_ms([&] { k_copy_ldg<<<sms * mult, 256>>>((const double2*)a, (double2*)b, n2); }, 9);
            printf(", \"ldg128_x%d\": %.1f", mult, 2.0 * bytes / ms / 1e6);
        }
        for (int mult : {2, 4, 8}) {
            ms = time_ms([&] { k_copy_ldg4<<<sms * mult, 256>>>((const double2*)a, (double2*)b, n2); }, 9);
            printf(", \"ldg128_ilp4_x%d\": %.1f", mult, 2.0 * bytes / ms / 1e6);
        }
        {
            constexpr int TILE = 16384, ST = 6;
            CK(cudaFuncSetAttribute(k_copy_tma<TILE, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * ST));
            for (int mult : {1, 2}) {
                ms = time_ms([&] { k_copy_tma<TILE, ST><<<sms * mult, 128, TILE * ST>>>(a, b, bytes / TILE); }, 9);
                printf(", \"tma16k_s6_x%d\": %.1f", mult, 2.0 * bytes / ms / 1e6);
            }
        }
        {
            constexpr int TILE = 32768, ST = 3;
            CK(cudaFuncSetAttribute(k_copy_tma<TILE, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * ST));
            for (int mult : {1, 2}) {
                ms = time_ms([&] { k_copy_tma<TILE, ST><<<sms * mult, 128, TILE * ST>>>(a, b, bytes / TILE); }, 9);
                printf(", \"tma32k_s3_x%d\": %.1f", mult, 2.0 * bytes / ms / 1e6);
            }
        }
        {
            constexpr int TILE = 8192, ST = 8;
            CK(cudaFuncSetAttribute(k_copy_tma<TILE, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * ST));
            for (int mult : {2, 3}) {
                ms = time_ms([&] { k_copy_tma<TILE, ST><<<sms * mult, 128, TILE * ST>>>(a, b, bytes / TILE); }, 9);
                printf(", \"tma8k_s8_x%d\": %.1f", mult, 2.0 * bytes / ms / 1e6);
            }
        }
        // verify the TMA copy actually copied
        CK(cudaMemset(b, 0, bytes));
        k_copy_tma<8192, 8><<<sms * 2, 128, 8192 * 8>>>(a, b, bytes / 8192);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned char> h(4096);
        CK(cudaMemcpy(h.data(), b + bytes - 4096, 4096, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (auto v : h) bad += (v != 1);
        printf(", \"tma_copy_bad_bytes\": %d}", bad);
        cudaFree(a), cudaFree(b);
    }
    printf("}\n");
    return 0;
}
