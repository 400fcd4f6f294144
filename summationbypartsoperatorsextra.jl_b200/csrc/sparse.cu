// sparse.cu -- K2/K3: batched apply of a structurally sparse operator (banded FSBP, ellipsoid-neighbourhood
// MFSBP) with the 2D-advection SAT term fused into the same pass.
//
// Reference math (src/conservation_laws/multidimensional_linear_advection.jl:73-81 with set_bc! :60-71):
//     du = -A u + B (u - tmp1)     =>  du_j = -sum_k A_jk u_k + [mode_j == inflow] * b_j * (u_j - g_j)
// and plain  y = alpha*scale*D u + beta*y  for mul! on operators stored dense-but-sparse.
//
// Layout: a CTA stages a tile of T instances into shared memory TRANSPOSED ([node][T], 16-byte chunks
// XOR-swizzled by the node index) so that one thread, owning one operator row, reads the T values of a
// column with 128-bit LDS and amortises each (col,val) pair over T FMAs.  The operator is stored in
// SELL-32 slices (32 rows, entry j of all 32 rows contiguous) so its loads are coalesced; it is shared by
// all CTAs and stays L2 resident.  Outputs go straight to global memory: for a fixed instance the 32 lanes of
// a warp own 32 consecutive nodes -> 256-byte coalesced stores.
#include <algorithm>
#include "common.cuh"
#include "ptx.cuh"

namespace sbp {

struct SparseParams {
    const double* __restrict__ U;
    double* __restrict__ dU;
    const int* __restrict__ slice_ptr;
    const int* __restrict__ sell_col;
    const double* __restrict__ sell_val;
    int n_slices;
    int N;
    long long B;
    double alpha;
    double beta;
    const double* __restrict__ colscale;  // a .* u (1D advection) or nullptr
    // fused SAT (2D advection) -- all nullptr for a plain apply
    const int* __restrict__ mode;
    const double* __restrict__ bdiag;
    const double* __restrict__ g;
    long long g_stride;
};

template <int T>
__device__ __forceinline__ int swz_off(int node, int t) {
    // doubles offset inside the [node][T] tile; 16-byte chunk index XORed with bits of the node index
    constexpr int CH = T / 2;  // chunks per node
    const int chunk = (t >> 1) ^ ((node >> 1) & (CH - 1) & 3);
    return node * T + chunk * 2 + (t & 1);
}

template <int T, int THREADS, bool SMEM>
__global__ void __launch_bounds__(THREADS) k_sparse(const SparseParams p) {
    extern __shared__ __align__(16) double s_u[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int WARPS = THREADS / 32;
    constexpr int CH = T / 2;
    const int N = p.N;
    const long long ntiles = (p.B + T - 1) / T;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long b0 = tile * T;
        const int nb = (int)std::min<long long>(T, p.B - b0);
        if (SMEM) {
            __syncthreads();  // previous tile fully consumed
            // ---- stage + transpose: coalesced 256-byte reads per (instance, 32 nodes) ----
            for (int t = warp; t < T; t += WARPS) {
                const double* src = p.U + (b0 + t) * N;
                const bool valid = t < nb;
                for (int i = lane; i < N; i += 32) {
                    double v = valid ? ldg_stream1(src + i) : 0.0;
                    if (p.colscale) v *= p.colscale[i];
                    s_u[swz_off<T>(i, t)] = v;
                }
            }
            __syncthreads();
        }
        // ---- one operator row per lane, SELL-32 slices round-robin over warps ----
        for (int s = warp; s < p.n_slices; s += WARPS) {
            const int row = s * 32 + lane;
            const int beg = p.slice_ptr[s], end = p.slice_ptr[s + 1];
            double acc[T];
#pragma unroll
            for (int t = 0; t < T; t++) acc[t] = 0.0;
            for (int j = beg; j < end; j++) {
                const int c = __ldg(p.sell_col + (size_t)j * 32 + lane);
                const double v = __ldg(p.sell_val + (size_t)j * 32 + lane);
                if (SMEM) {
                    const int sw = (c >> 1) & (CH - 1) & 3;
                    const double* base = s_u + (size_t)c * T;
#pragma unroll
                    for (int lc = 0; lc < CH; lc++) {
                        const double2 u2 = *reinterpret_cast<const double2*>(base + ((lc ^ sw) << 1));
                        acc[2 * lc] = fma(v, u2.x, acc[2 * lc]);
                        acc[2 * lc + 1] = fma(v, u2.y, acc[2 * lc + 1]);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < T; t++) {
                        if (t < nb) {
                            double u = __ldg(p.U + (b0 + t) * N + c);
                            if (p.colscale) u *= p.colscale[c];
                            acc[t] = fma(v, u, acc[t]);
                        }
                    }
                }
            }
            if (row < N) {
                double sat_b = 0.0;
                bool inflow = false;
                if (p.mode) {
                    inflow = p.mode[row] == 1;
                    sat_b = p.bdiag[row];
                }
#pragma unroll
                for (int t = 0; t < T; t++) {
                    if (t < nb) {
                        double y = p.alpha * acc[t];
                        if (inflow) {
                            // the SAT uses the unscaled u_j (colscale is never combined with a 2D SAT)
                            const double uj = SMEM ? s_u[swz_off<T>(row, t)] : p.U[(b0 + t) * N + row];
                            const double gj = p.g[(b0 + t) * p.g_stride + row];
                            y += sat_b * (uj - gj);
                        }
                        double* yp = p.dU + (b0 + t) * N + row;
                        if (p.beta != 0.0) y = fma(p.beta, *yp, y);
                        stg_stream1(yp, y);
                    }
                }
            }
        }
    }
}

template <int T, bool SMEM>
static int32_t launch_sparse_t(sbp_ctx* ctx, const SparseParams& p) {
    constexpr int THREADS = 256;
    auto kern = k_sparse<T, THREADS, SMEM>;
    size_t smem = SMEM ? (size_t)p.N * T * sizeof(double) : 0;
    if (SMEM) SBP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SBP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const long long ntiles = (p.B + T - 1) / T;
    int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)ctx->sm_count * per_sm));
    kern<<<grid, THREADS, smem, ctx->stream>>>(p);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

int32_t launch_sparse_ex(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                         double beta, const sbp_op* adv, const double* g, int64_t g_stride,
                         const double* colscale) {
    SparseParams p;
    p.U = U;
    p.dU = dU;
    p.slice_ptr = op->d_slice_ptr;
    p.sell_col = op->d_sell_col;
    p.sell_val = op->d_sell_val;
    p.n_slices = op->n_slices;
    p.N = op->N;
    p.B = B;
    p.alpha = alpha * op->scale;
    p.beta = beta;
    p.colscale = colscale;
    p.mode = adv ? adv->d_mode : nullptr;
    p.bdiag = adv ? adv->d_b : nullptr;
    p.g = g;
    p.g_stride = g_stride;
    const size_t budget = (size_t)ctx->smem_optin - 2048;
    const size_t half = budget / 2 - 1024;  // two CTAs per SM overlap staging and compute
    const size_t n8 = (size_t)op->N * 8;
    if (n8 * 16 <= half) return launch_sparse_t<16, true>(ctx, p);
    if (n8 * 8 <= half) return launch_sparse_t<8, true>(ctx, p);
    if (n8 * 8 <= budget) return launch_sparse_t<8, true>(ctx, p);
    if (n8 * 4 <= budget) return launch_sparse_t<4, true>(ctx, p);
    return launch_sparse_t<4, false>(ctx, p);  // vectors too long for shared memory: gather through L1/L2
}

int32_t launch_sparse(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                      double beta, const sbp_op* adv, const double* g, int64_t g_stride, double* quantities) {
    int32_t rc = launch_sparse_ex(ctx, op, dU, U, B, alpha, beta, adv, g, g_stride, nullptr);
    if (rc) return rc;
    if (adv && quantities) return launch_adv2d_epilogue(ctx, adv, dU, U, B, g, g_stride, quantities, false);
    return SBP_OK;
}

}  // namespace sbp
