// table_generic.cu -- K4 fallback: operator-table apply with an ARBITRARY per-instance operator index
// (the periodic assignment b % G takes the DMMA path in dense_small.cu).  One warp per instance, lanes own
// output rows, the instance vector is staged in shared memory, operators are read column-major so that the
// 32 lanes of a warp read 32 consecutive doubles; the G small operators stay L1/L2 resident.
#include <algorithm>
#include "common.cuh"
#include "ptx.cuh"

namespace sbp {

__global__ void __launch_bounds__(256) k_table_generic(const double* __restrict__ U, double* __restrict__ dU,
                                                       const double* __restrict__ Dcm, const double* __restrict__ w,
                                                       int N, int G, long long B, const int* __restrict__ op_index,
                                                       double alpha, double beta, double* __restrict__ energy,
                                                       double* __restrict__ partial) {
    extern __shared__ double s_u[];  // 8 warps x N
    __shared__ double warp_part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* su = s_u + warp * N;
    double e_acc = 0.0;
    for (long long b = (long long)blockIdx.x * 8 + warp; b < B; b += (long long)gridDim.x * 8) {
        int gi = op_index ? op_index[b] : (int)(b % G);
        gi = gi < 0 ? 0 : (gi >= G ? G - 1 : gi);
        const double* u = U + b * N;
        const double* D = Dcm + (size_t)gi * N * N;
        double e = 0.0;
        __syncwarp();
        for (int i = lane; i < N; i += 32) {
            const double ui = ldg_stream1(u + i);
            su[i] = ui;
            if (w) e = fma(w[(size_t)gi * N + i] * ui, ui, e);
        }
        __syncwarp();
        if (energy || partial) {
            e = warp_sum(e);
            if (lane == 0) {
                if (energy) energy[b] = e;
                e_acc += e;
            }
        }
        for (int i = lane; i < N; i += 32) {
            double acc = 0.0;
            for (int k = 0; k < N; k++) acc = fma(__ldg(D + (size_t)k * N + i), su[k], acc);
            double y = alpha * acc;
            double* yp = dU + b * N + i;
            if (beta != 0.0) y = fma(beta, *yp, y);
            stg_stream1(yp, y);
        }
    }
    if (partial) {
        if (lane == 0) warp_part[warp] = e_acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int i = 0; i < 8; i++) s += warp_part[i];
            partial[blockIdx.x] = s;
        }
    }
}

int32_t launch_table_generic(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B,
                             const int32_t* op_index, double alpha, double beta, double* energy,
                             double* energy_sum) {
    const bool want_energy = energy || energy_sum;
    SBP_REQUIRE(!want_energy || op->d_w, SBP_EINVAL, "energy requested but the operator has no weights");
    const size_t smem = (size_t)8 * op->N * sizeof(double);
    int grid = (int)std::max<long long>(1, std::min<long long>((B + 7) / 8, (long long)ctx->sm_count * 8));
    double* partial = nullptr;
    if (energy_sum) {
        int32_t rc = ensure_scratch(ctx, grid);
        if (rc) return rc;
        partial = ctx->scratch;
    }
    SBP_CUDA(cudaFuncSetAttribute(k_table_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // d_D holds the table ROW-major; the column-major copy lives right behind it (see api.cu)
    const double* Dcm = op->d_D + (size_t)op->G * op->N * op->N;
    k_table_generic<<<grid, 256, smem, ctx->stream>>>(U, dU, Dcm, want_energy ? op->d_w : nullptr, op->N, op->G, B,
                                                      op_index, alpha * op->scale, beta, energy, partial);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    if (energy_sum) return launch_sum_partials(ctx, partial, grid, energy_sum);
    return SBP_OK;
}

}  // namespace sbp
