// dense_gemm.cu -- K6: batched apply of a LARGE dense operator (N > 128, e.g. the dense MFSBP operator
// N = 4096 of BASELINE config 5): a genuine FP64 GEMM
//        Y[inst][i] = alpha' * sum_k U[inst][k] * D[i][k] + beta * Y[inst][i]
// with instances on the M axis.  Both operands are K-contiguous in memory (U rows are instances, D is
// kept row-major on the device), so one multi-stage cp.async pipeline feeds padded shared-memory tiles and
// every warp issues DMMA.8x8x4 on 8x4 / 4x8 fragments read with conflict-free 64-bit LDS.
// tcgen05.mma has no f64 kind: on sm_100a the FP64 tensor path is mma.sync (SASS DMMA).
#include <algorithm>
#include "common.cuh"
#include "ptx.cuh"

namespace sbp {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// CTA tile BM x BN x BK, WGM x WGN warps, each warp (BM/WGM) x (BN/WGN); LD = BK + 4 keeps LD % 16 == 4, which makes the
// 8-row x 4-column fragment loads (lane = 4*row + col) hit 32 distinct 8-byte banks.
template <int BM_, int BN_, int BK_, int WGM_, int WGN_, int STAGES_>
struct GemmCfg {
    static constexpr int BM = BM_, BN = BN_, BK = BK_, WGM = WGM_, WGN = WGN_, STAGES = STAGES_;
    static constexpr int LD = BK + 4;
    static constexpr int THREADS = WGM * WGN * 32;
    static constexpr int WM = BM / WGM, WN = BN / WGN;
    static constexpr int MT = WM / 8, NTW = WN / 8;
    static constexpr int STAGE_DOUBLES = (BM + BN) * LD;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_DOUBLES * sizeof(double);
};

template <class C, bool VEC, int ROWS>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ G, long long row0, long long rows,
                                          int k0, int K, int tid) {
    constexpr int CHUNKS_PER_ROW = C::BK / 2;  // 16-byte chunks
    constexpr int TOTAL = ROWS * CHUNKS_PER_ROW;
#pragma unroll
    for (int it = 0; it < (TOTAL + C::THREADS - 1) / C::THREADS; it++) {
        const int chunk = tid + it * C::THREADS;
        if (TOTAL % C::THREADS != 0 && chunk >= TOTAL) break;
        const int r = chunk / CHUNKS_PER_ROW, c = (chunk % CHUNKS_PER_ROW) * 2;
        const long long gr = row0 + r;
        const int gk = k0 + c;
        double* dst = s + r * C::LD + c;
        if (VEC) {
            int bytes = 0;
            if (gr < rows && gk < K) bytes = (K - gk >= 2) ? 16 : 8;
            const double* src = (bytes > 0) ? G + gr * K + gk : G;
            cp_async16(dst, src, bytes);
        } else {
            dst[0] = (gr < rows && gk < K) ? G[gr * K + gk] : 0.0;
            dst[1] = (gr < rows && gk + 1 < K) ? G[gr * K + gk + 1] : 0.0;
        }
    }
}

template <class C, bool VEC>
__global__ void __launch_bounds__(C::THREADS, 1)
    k_dense_gemm(const double* __restrict__ U, const double* __restrict__ Drm, double* __restrict__ Y, long long B,
                 int N, double alpha, double beta, int tiles_n, long long tiles_total) {
    constexpr int BM = C::BM, BN = C::BN, BK = C::BK, LD = C::LD, STAGES = C::STAGES, MT = C::MT, NTW = C::NTW;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp / C::WGN) * C::WM, wn = (warp % C::WGN) * C::WN;
    const int KT = (N + BK - 1) / BK;

    for (long long tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
        // n fastest: CTAs resident together share the same rows of U through L2
        const long long tm = tile / tiles_n;
        const int tn = (int)(tile % tiles_n);
        const long long m0 = tm * BM;
        const int n0 = tn * BN;

        double c[MT][NTW][2];
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NTW; j++) c[i][j][0] = c[i][j][1] = 0.0;

        __syncthreads();  // smem free (previous tile done)
#pragma unroll
        for (int s = 0; s < STAGES - 1; s++) {
            if (s < KT) {
                double* sa = smem + s * C::STAGE_DOUBLES;
                load_tile<C, VEC, BM>(sa, U, m0, B, s * BK, N, tid);
                load_tile<C, VEC, BN>(sa + BM * LD, Drm, n0, N, s * BK, N, tid);
            }
            cp_async_commit();
        }
        for (int kt = 0; kt < KT; kt++) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            {   // prefetch stage kt+STAGES-1 into the slot freed at iteration kt-1
                const int kn = kt + STAGES - 1;
                if (kn < KT) {
                    double* sa = smem + (kn % STAGES) * C::STAGE_DOUBLES;
                    load_tile<C, VEC, BM>(sa, U, m0, B, kn * BK, N, tid);
                    load_tile<C, VEC, BN>(sa + BM * LD, Drm, n0, N, kn * BK, N, tid);
                }
                cp_async_commit();
            }
            const double* sa = smem + (kt % STAGES) * C::STAGE_DOUBLES;
            const double* sb = sa + BM * LD;
#pragma unroll
            for (int k4 = 0; k4 < BK / 4; k4++) {
                double af[MT], bf[NTW];
#pragma unroll
                for (int i = 0; i < MT; i++) af[i] = sa[(wm + i * 8 + g) * LD + k4 * 4 + t];
#pragma unroll
                for (int j = 0; j < NTW; j++) bf[j] = sb[(wn + j * 8 + g) * LD + k4 * 4 + t];
#pragma unroll
                for (int i = 0; i < MT; i++)
#pragma unroll
                    for (int j = 0; j < NTW; j++) dmma884(c[i][j][0], c[i][j][1], af[i], bf[j]);
            }
        }
        cp_async_wait<0>();
        // ---- epilogue ----
#pragma unroll
        for (int i = 0; i < MT; i++) {
            const long long m = m0 + wm + i * 8 + g;
            if (m >= B) continue;
#pragma unroll
            for (int j = 0; j < NTW; j++) {
                const int n = n0 + wn + j * 8 + 2 * t;
                double y0 = alpha * c[i][j][0], y1 = alpha * c[i][j][1];
                double* yp = Y + m * N + n;
                if (VEC && n + 1 < N) {
                    if (beta != 0.0) {
                        const double2 o = *reinterpret_cast<const double2*>(yp);
                        y0 = fma(beta, o.x, y0);
                        y1 = fma(beta, o.y, y1);
                    }
                    *reinterpret_cast<double2*>(yp) = make_double2(y0, y1);
                } else {
                    if (n < N) yp[0] = (beta != 0.0) ? fma(beta, yp[0], y0) : y0;
                    if (n + 1 < N) yp[1] = (beta != 0.0) ? fma(beta, yp[1], y1) : y1;
                }
            }
        }
    }
}

template <class C>
static int32_t launch_cfg(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                          double beta) {
    const int N = op->N;
    const int tiles_n = (N + C::BN - 1) / C::BN;
    const long long tiles_m = (B + C::BM - 1) / C::BM;
    const long long total = tiles_m * tiles_n;
    const bool vec = (N % 2 == 0) && ((reinterpret_cast<uintptr_t>(U) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(dU) & 15) == 0);
    const int grid = (int)std::max<long long>(1, std::min<long long>(total, ctx->sm_count));
    if (vec) {
        SBP_CUDA(cudaFuncSetAttribute(k_dense_gemm<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        k_dense_gemm<C, true><<<grid, C::THREADS, C::SMEM, ctx->stream>>>(U, op->d_D, dU, B, N, alpha * op->scale, beta,
                                                                          tiles_n, total);
    } else {
        SBP_CUDA(cudaFuncSetAttribute(k_dense_gemm<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        k_dense_gemm<C, false><<<grid, C::THREADS, C::SMEM, ctx->stream>>>(U, op->d_D, dU, B, N, alpha * op->scale,
                                                                           beta, tiles_n, total);
    }
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

int32_t launch_dense_gemm(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                          double beta) {
    // tuning knob SBP_GEMM_VARIANT (tools/time_config.py --workload config5 sweeps it)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SBP_GEMM_VARIANT");
        v = e ? atoi(e) : 0;
    }
    switch (v) {
        case 1: return launch_cfg<GemmCfg<128, 128, 16, 4, 4, 4>>(ctx, op, dU, U, B, alpha, beta);  // 16 warps, 32x32
        case 2: return launch_cfg<GemmCfg<128, 128, 32, 2, 4, 3>>(ctx, op, dU, U, B, alpha, beta);  // BK = 32
        case 3: return launch_cfg<GemmCfg<128, 128, 32, 4, 4, 3>>(ctx, op, dU, U, B, alpha, beta);
        case 4: return launch_cfg<GemmCfg<128, 128, 16, 2, 4, 3>>(ctx, op, dU, U, B, alpha, beta);  // 3 stages
        case 5: return launch_cfg<GemmCfg<128, 128, 16, 4, 2, 4>>(ctx, op, dU, U, B, alpha, beta);  // 32x64 warp tiles
        default: return launch_cfg<GemmCfg<128, 128, 16, 2, 4, 4>>(ctx, op, dU, U, B, alpha, beta);  // 8 warps, 64x32
    }
}

}  // namespace sbp
