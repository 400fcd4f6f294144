// dense_gemm.cu -- K6: batched apply of a LARGE dense operator (N > 128, e.g. the dense MFSBP operator
// N = 4096 of BASELINE config 5): a genuine FP64 GEMM
//        Y[inst][i] = alpha' * sum_k U[inst][k] * D[i][k] + beta * Y[inst][i]
// with instances on the M axis.  Both operands are K-contiguous in memory (U rows are instances, D is
// kept row-major on the device), so one multi-stage cp.async pipeline feeds padded shared-memory tiles and
// every warp issues DMMA.8x8x4 on 8x4 / 4x8 fragments read with conflict-free 64-bit LDS.
// tcgen05.mma has no f64 kind: on sm_100a the FP64 tensor path is mma.sync (SASS DMMA).
#include <algorithm>
#include <cuda.h>
#include "common.cuh"
#include "ptx.cuh"

namespace sbp {

int32_t make_tmap(CUtensorMap* tm, const double* base, int N, long long B, int box_nodes, int box_inst);  // bsr.cu

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


// ---------------------------------------------------------------------------------------------------------------
// K6-TMA: the same GEMM with a warp-specialised TMA/mbarrier pipeline.  One producer thread streams {BK x 128} boxes of
// U and D (cp.async.bulk.tensor.2d) through a ring of stages guarded by full/empty mbarriers; eight consumer warps
// (2 x 4, warp tile 64 x 32) do nothing but 64-bit LDS + DMMA.8x8x4: no cp.async address arithmetic in the math warps,
// no CTA-wide barrier in the main loop, and the next tile's first stages stream in during the epilogue.
// BK = 20 doubles: a box row is 160 bytes = 4 (mod 16) doubles, so the 8-row x 4-column fragment gathers are
// bank-conflict free without padding (which a TMA box cannot have) and without swizzling (SWIZZLE_128B leaves 2-way
// conflicts for 8-byte elements in this fragment shape).  K % 20 != 0 is handled by the TMA's zero fill.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GT_BM = 128, GT_BN = 128, GT_BK = 20, GT_ST = 5, GT_CONS = 8;
constexpr int GT_STAGE = (GT_BM + GT_BN) * GT_BK;  // doubles per stage
constexpr size_t GT_SMEM = (size_t)GT_ST * GT_STAGE * sizeof(double);

__device__ __forceinline__ void gt_tma_load(void* smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
            "r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

__global__ void __launch_bounds__((GT_CONS + 1) * 32, 1)
    k_dense_gemm_tma(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmD,
                     double* __restrict__ Y, long long B, int N, double alpha, double beta, int tiles_n,
                     long long tiles_total) {
    constexpr int BM = GT_BM, BN = GT_BN, BK = GT_BK, ST = GT_ST;
    constexpr int WM = 64, WN = 32, MT = WM / 8, NTW = WN / 8, KS = BK / 4;
    extern __shared__ __align__(128) double smem[];
    __shared__ uint64_t bar_full[ST], bar_empty[ST];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int KT = (N + BK - 1) / BK;
    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; s++) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], GT_CONS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == GT_CONS) {  // ---- producer ----
        if (lane == 0) {
            int st = 0, round = 0;
            for (long long tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
                const int m0 = (int)(tile / tiles_n) * BM, n0 = (int)(tile % tiles_n) * BN;
                for (int kt = 0; kt < KT; kt++) {
                    if (round > 0) mbar_wait(&bar_empty[st], (round - 1) & 1);
                    double* sa = smem + (size_t)st * GT_STAGE;
                    mbar_arrive_expect_tx(&bar_full[st], GT_STAGE * 8);
                    gt_tma_load(sa, &tmU, kt * BK, m0, &bar_full[st]);
                    gt_tma_load(sa + BM * BK, &tmD, kt * BK, n0, &bar_full[st]);
                    if (++st == ST) st = 0, round++;
                }
            }
        }
        return;
    }

    // ---- consumers ----
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp / 4) * WM, wn = (warp % 4) * WN;
    const int aoff = (wm + g) * BK + t, boff = BM * BK + (wn + g) * BK + t;
    int st = 0, round = 0;
    for (long long tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
        const long long m0 = (tile / tiles_n) * BM;
        const int n0 = (int)(tile % tiles_n) * BN;
        double c[MT][NTW][2];
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NTW; j++) c[i][j][0] = c[i][j][1] = 0.0;

        for (int kt = 0; kt < KT; kt++) {
            mbar_wait(&bar_full[st], round & 1);
            const double* sa = smem + (size_t)st * GT_STAGE + aoff;
            const double* sb = smem + (size_t)st * GT_STAGE + boff;
            double af[2][MT], bf[2][NTW];
#pragma unroll
            for (int i = 0; i < MT; i++) af[0][i] = sa[i * 8 * BK];
#pragma unroll
            for (int j = 0; j < NTW; j++) bf[0][j] = sb[j * 8 * BK];
#pragma unroll
            for (int k4 = 0; k4 < KS; k4++) {
                const int cur = k4 & 1, nxt = cur ^ 1;
                if (k4 + 1 < KS) {  // fragments of the next k-step are in flight under this step's DMMA chain
#pragma unroll
                    for (int i = 0; i < MT; i++) af[nxt][i] = sa[i * 8 * BK + (k4 + 1) * 4];
#pragma unroll
                    for (int j = 0; j < NTW; j++) bf[nxt][j] = sb[j * 8 * BK + (k4 + 1) * 4];
                }
#pragma unroll
                for (int i = 0; i < MT; i++)
#pragma unroll
                    for (int j = 0; j < NTW; j++) dmma884(c[i][j][0], c[i][j][1], af[cur][i], bf[cur][j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[st]);
            if (++st == ST) st = 0, round++;
        }
        // ---- epilogue: lane (g,t) owns C[8i+g][8j+2t..2t+1]; 4 lanes write 64 contiguous bytes ----
#pragma unroll
        for (int i = 0; i < MT; i++) {
            const long long m = m0 + wm + i * 8 + g;
            if (m >= B) continue;
#pragma unroll
            for (int j = 0; j < NTW; j++) {
                const int n = n0 + wn + j * 8 + 2 * t;
                if (n >= N) continue;  // N is even on this path: n + 1 < N
                double y0 = alpha * c[i][j][0], y1 = alpha * c[i][j][1];
                double* yp = Y + m * N + n;
                if (beta != 0.0) {
                    const double2 o = *reinterpret_cast<const double2*>(yp);
                    y0 = fma(beta, o.x, y0);
                    y1 = fma(beta, o.y, y1);
                }
                *reinterpret_cast<double2*>(yp) = make_double2(y0, y1);
            }
        }
    }
}

static int32_t launch_gemm_tma(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                               double beta) {
    const int N = op->N;
    const int tiles_n = (N + GT_BN - 1) / GT_BN;
    const long long tiles_m = (B + GT_BM - 1) / GT_BM;
    const long long total = tiles_m * tiles_n;
    CUtensorMap tmU, tmD;
    int32_t rc = make_tmap(&tmU, U, N, B, GT_BK, GT_BM);
    if (!rc) rc = make_tmap(&tmD, op->d_D, N, N, GT_BK, GT_BN);
    if (rc) return rc;
    const int grid = (int)std::max<long long>(1, std::min<long long>(total, ctx->sm_count));
    SBP_CUDA(cudaFuncSetAttribute(k_dense_gemm_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM));
    k_dense_gemm_tma<<<grid, (GT_CONS + 1) * 32, GT_SMEM, ctx->stream>>>(tmU, tmD, dU, B, N, alpha * op->scale, beta,
                                                                         tiles_n, total);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

int32_t launch_dense_gemm(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                          double beta) {
    // tuning knob SBP_GEMM_VARIANT (tools/time_config.py --workload config5 sweeps it)
    const int v = ctx->tune[T_GEMM_VARIANT];
    // default: the TMA pipeline (needs 16-byte aligned rows: N even, aligned base pointers, B < 2^31); SBP_GEMM_VARIANT
    // >= 0 selects one of the cp.async configurations, which also serve the unaligned cases
    const bool tma_ok = (op->N % 2 == 0) && ((reinterpret_cast<uintptr_t>(U) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(dU) & 15) == 0) && B < (1ll << 31) - 256;
    if (v < 0 && tma_ok) return launch_gemm_tma(ctx, op, dU, U, B, alpha, beta);
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
