// sparse.cu -- K2/K3: batched apply of a structurally sparse operator (banded FSBP, ellipsoid-neighbourhood
// MFSBP) with the 2D-advection SAT term fused into the same pass.
//
// Reference math (src/conservation_laws/multidimensional_linear_advection.jl:73-81 with set_bc! :60-71):
//     du = -A u + B (u - tmp1)     =>  du_j = -sum_k A_jk u_k + [mode_j == inflow] * b_j * (u_j - g_j)
// and plain  y = alpha*scale*D u + beta*y  for mul! on operators stored dense-but-sparse.
//
// Design (the gather from shared memory, 8 B per multiply-add, is what bounds this kernel -- not HBM):
//  * a CTA stages a tile of T = 8 instances into shared memory TRANSPOSED ([node slot][8 instances], 64 B per
//    node) so one thread reads the 8 values of a column with four 128-bit LDS and amortises each operator entry
//    over 8 FMAs;
//  * ROW BLOCKS: a thread owns R (1, 2 or 4, chosen at upload) consecutive rows and walks the UNION of their
//    column sets; every gathered column feeds R x 8 FMAs from registers.  For stencil-like operators
//    neighbouring rows share most columns, which cuts the shared-memory traffic per output by up to ~2x;
//    values are stored zero-padded per union column (SELL-32 slices over row blocks => coalesced operator
//    loads, L2 resident);
//  * node -> slot permutation + XOR swizzle of the 16-byte chunks make both the staging stores and the
//    gathers of uniformly strided column patterns bank-conflict free (slot parity = row-block parity);
//  * a lane's R rows are contiguous in the output, lanes are contiguous row blocks: 64/128/256-bit stores of a
//    warp cover one contiguous span per instance.
#include <algorithm>
#include "common.cuh"
#include "ptx.cuh"

namespace sbp {

struct SparseParams {
    const double* __restrict__ U;
    double* __restrict__ dU;
    const int* __restrict__ slice_ptr;   // n_slices + 1 (entries per slice, prefix sum)
    const int* __restrict__ cols;        // [entry][32] slot index of the union column
    const double* __restrict__ vals;     // [entry][R][32]
    int n_slices;
    int N;
    int n_slots;                         // N rounded up to a multiple of 2R
    long long B;
    double alpha;
    double beta;
    const double* __restrict__ colscale;  // a .* u (1D advection) or nullptr
    // fused SAT (2D advection) -- all nullptr for a plain apply
    const int* __restrict__ mode;
    const double* __restrict__ bdiag;
    const double* __restrict__ g;
    long long g_stride;
    int vec_ok;                          // output rows of a lane may be stored with one vector store
    Sat1D sat;                           // fused 1D advection SATs (rows 0 and N-1); the staged tile holds a .* u
};

constexpr int kT = 8;  // instances per tile

template <int LOGR>
__host__ __device__ __forceinline__ int node_slot(int c) {
    constexpr int R = 1 << LOGR;
    const int cb = c >> LOGR, r = c & (R - 1);
    return ((cb >> 1) << (LOGR + 1)) | (r << 1) | (cb & 1);
}
template <int LOGR>
__device__ __forceinline__ int slot_swz(int slot) {
    return (slot >> (LOGR + 1)) & 3;
}

__device__ __forceinline__ void stg256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

constexpr int kSparseMaxThreads = 384;

// Stage + transpose one tile of kT instances into `s_tile` ([slot][kT], swizzled).  A warp handles 32 consecutive nodes
// for all kT instances: coalesced 256-byte reads; lanes are permuted so that a quarter-warp writes 8 different row
// blocks (no bank conflicts).  Without column scaling the copies are asynchronous LDGSTS (cp.async, 8 bytes each): the
// caller waits with cp_async_wait_all() + __syncthreads() -- possibly a whole tile of compute later (double buffering).
template <int LOGR>
__device__ __forceinline__ void stage_tile(const SparseParams& p, double* s_tile, long long tile, int lane, int warp,
                                           int WARPS) {
    constexpr int R = 1 << LOGR;
    const int N = p.N;
    const long long b0 = tile * kT;
    const int nb = (p.B - b0) < kT ? (int)(p.B - b0) : kT;
    const int lperm = (lane % (32 / R)) * R + lane / (32 / R);
    const int nchunks = (p.n_slots + 31) / 32;
    for (int chunk = warp; chunk < nchunks; chunk += WARPS) {
        const int node = chunk * 32 + lperm;
        if (node >= p.n_slots) continue;
        const int slot = node_slot<LOGR>(node);
        const int sw = slot_swz<LOGR>(slot);
        double* dst = s_tile + slot * kT;
        const double* src = p.U + b0 * N + node;
        if (p.colscale == nullptr) {
#pragma unroll
            for (int t = 0; t < kT; t++) {
                const bool ok = node < N && t < nb;
                cp_async8(dst + (((t >> 1) ^ sw) << 1) + (t & 1), ok ? src + (long long)t * N : p.U, ok ? 8 : 0);
            }
        } else {
            double v[kT];
#pragma unroll
            for (int t = 0; t < kT; t++) v[t] = (node < N && t < nb) ? ldg_stream1(src + (long long)t * N) : 0.0;
            const double cs = node < N ? p.colscale[node] : 0.0;
#pragma unroll
            for (int q = 0; q < kT / 2; q++)
                *reinterpret_cast<double2*>(dst + ((q ^ sw) << 1)) = make_double2(v[2 * q] * cs, v[2 * q + 1] * cs);
        }
    }
}

// DB: two tile buffers -- the next tile is staged (asynchronously) while the current one is being consumed.
template <int LOGR, bool SMEM, bool DB>
__global__ void __launch_bounds__(kSparseMaxThreads) k_sparse(const SparseParams p) {
    constexpr int R = 1 << LOGR;
    const int WARPS = blockDim.x >> 5;
    extern __shared__ __align__(16) double s_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.N;
    const long long ntiles = (p.B + kT - 1) / kT;
    const size_t tile_doubles = (size_t)p.n_slots * kT;
    int buf = 0;
    if (SMEM && DB && (long long)blockIdx.x < ntiles) {
        stage_tile<LOGR>(p, s_all, blockIdx.x, lane, warp, WARPS);
        cp_async_wait_all();
        __syncthreads();
    }

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long b0 = tile * kT;
        const int nb = (p.B - b0) < kT ? (int)(p.B - b0) : kT;
        double* s_u = s_all + (DB ? buf * tile_doubles : 0);
        if (SMEM && !DB) {
            __syncthreads();  // previous tile fully consumed
            stage_tile<LOGR>(p, s_u, tile, lane, warp, WARPS);
            cp_async_wait_all();
            __syncthreads();
        }
        if (SMEM && DB && tile + gridDim.x < ntiles)
            stage_tile<LOGR>(p, s_all + (buf ^ 1) * tile_doubles, tile + gridDim.x, lane, warp, WARPS);
        // ---- one row block (R rows) per lane, SELL-32 slices of row blocks round-robin over warps ----
        for (int s = warp; s < p.n_slices; s += WARPS) {
            const int rb = s * 32 + lane;
            const int beg = p.slice_ptr[s], end = p.slice_ptr[s + 1];
            // boundary slices of a 1D advection RHS: request the boundary data now, use it after the gather loop
            const bool has0 = (s == 0), hasN = (((N - 1) >> LOGR) >> 5) == s;  // warp-uniform
            double sat_bc = 0.0;
            if (p.sat.enabled && (has0 || hasN) && lane < 2 * kT && (lane & (kT - 1)) < nb)
                sat_bc = p.sat.bc[(b0 + (lane & (kT - 1))) * p.sat.bc_stride + ((lane >> 3) & 1)];
            double acc[R][kT];
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int t = 0; t < kT; t++) acc[r][t] = 0.0;
            const int* cp = p.cols + (size_t)beg * 32 + lane;
            const double* vp = p.vals + (size_t)beg * R * 32 + lane;
            // Slice widths are even (padded on the host): two union columns per step, software-pipelined one step
            // ahead so that the L2 latency of the operator entries overlaps the gathers of the previous step.
            int c0 = 0, c1 = 0;
            double v0[R], v1[R];
            if (beg < end) {
                c0 = __ldg(cp), c1 = __ldg(cp + 32);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    v0[r] = __ldg(vp + r * 32);
                    v1[r] = __ldg(vp + (R + r) * 32);
                }
            }
            for (int j = beg; j < end; j += 2) {
                const bool more = j + 2 < end;
                const int* cpn = cp + (more ? 64 : 0);
                const double* vpn = vp + (more ? 2 * R * 32 : 0);
                const int n0 = __ldg(cpn), n1 = __ldg(cpn + 32);
                double w0[R], w1[R];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    w0[r] = __ldg(vpn + r * 32);
                    w1[r] = __ldg(vpn + (R + r) * 32);
                }
                if (SMEM) {
                    const double* base0 = s_u + c0 * kT;
                    const double* base1 = s_u + c1 * kT;
                    const int sw0 = slot_swz<LOGR>(c0), sw1 = slot_swz<LOGR>(c1);
#pragma unroll
                    for (int q = 0; q < kT / 2; q++) {
                        const double2 a = *reinterpret_cast<const double2*>(base0 + ((q ^ sw0) << 1));
                        const double2 b = *reinterpret_cast<const double2*>(base1 + ((q ^ sw1) << 1));
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            acc[r][2 * q] = fma(v0[r], a.x, acc[r][2 * q]);
                            acc[r][2 * q + 1] = fma(v0[r], a.y, acc[r][2 * q + 1]);
                            acc[r][2 * q] = fma(v1[r], b.x, acc[r][2 * q]);
                            acc[r][2 * q + 1] = fma(v1[r], b.y, acc[r][2 * q + 1]);
                        }
                    }
                } else {
                    // vectors too long for shared memory: gather straight from global (cols hold NODE indices)
#pragma unroll
                    for (int t = 0; t < kT; t++) {
                        if (t < nb) {
                            double ua = __ldg(p.U + (b0 + t) * N + c0), ub = __ldg(p.U + (b0 + t) * N + c1);
                            if (p.colscale) {
                                ua *= p.colscale[c0];
                                ub *= p.colscale[c1];
                            }
#pragma unroll
                            for (int r = 0; r < R; r++) acc[r][t] = fma(v1[r], ub, fma(v0[r], ua, acc[r][t]));
                        }
                    }
                }
                c0 = n0, c1 = n1;
#pragma unroll
                for (int r = 0; r < R; r++) v0[r] = w0[r], v1[r] = w1[r];
                cp = cpn, vp = vpn;
            }
            // ---- fused 1D SATs (rows 0 and N-1): the slice(s) holding those rows compute the 2 x kT terms with 16 lanes in
            //      parallel (one boundary-data load and one division each) and hand them to the row owners by shuffle ----
            double sat_l[kT], sat_r[kT];
            if (p.sat.enabled && (has0 || hasN)) {
                const int tt = lane & (kT - 1), side = (lane >> 3) & 1;
                double satv = 0.0;
                if (lane < 2 * kT && tt < nb && (side ? hasN : has0)) {
                    const int row = side ? N - 1 : 0;
                    double tj;  // a_j * u_j exactly as the reference's a .* u (the staged tile is column scaled)
                    if (SMEM) {
                        const int slot = node_slot<LOGR>(row);
                        tj = s_u[slot * kT + (((tt >> 1) ^ slot_swz<LOGR>(slot)) << 1) + (tt & 1)];
                    } else {
                        tj = p.U[(b0 + tt) * N + row] * (p.colscale ? p.colscale[row] : 1.0);
                    }
                    satv = side ? sat1d_right_v(p.sat, sat_bc, tj) : sat1d_left_v(p.sat, sat_bc, tj);
                }
#pragma unroll
                for (int t = 0; t < kT; t++) {
                    sat_l[t] = __shfl_sync(0xffffffffu, satv, t);
                    sat_r[t] = __shfl_sync(0xffffffffu, satv, kT + t);
                }
            }
            // ---- epilogue: alpha, fused SAT on inflow rows, beta, vector store of the lane's R contiguous rows ----
            const int row0 = rb * R;
            if (row0 < N) {
                double satb[R];
                bool inflow[R];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    inflow[r] = false;
                    satb[r] = 0.0;
                    if (p.mode && row0 + r < N && p.mode[row0 + r] == 1) {
                        inflow[r] = true;
                        satb[r] = p.bdiag[row0 + r];
                    }
                }
#pragma unroll
                for (int t = 0; t < kT; t++) {
                    if (t >= nb) continue;
                    double y[R];
                    double* yp = p.dU + (b0 + t) * N + row0;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        y[r] = p.alpha * acc[r][t];
                        if (inflow[r]) {
                            // the SAT uses the unscaled u_j (colscale is never combined with a 2D SAT)
                            double uj;
                            if (SMEM) {
                                const int slot = node_slot<LOGR>(row0 + r);
                                uj = s_u[slot * kT + (((t >> 1) ^ slot_swz<LOGR>(slot)) << 1) + (t & 1)];
                            } else {
                                uj = p.U[(b0 + t) * N + row0 + r];
                            }
                            y[r] += satb[r] * (uj - p.g[(b0 + t) * p.g_stride + row0 + r]);
                        }
                        if (p.sat.enabled && (has0 || hasN)) {
                            if (row0 + r == 0) y[r] += sat_l[t];
                            if (row0 + r == N - 1) y[r] += sat_r[t];
                        }
                    }
                    const bool full = row0 + R <= N;
                    if (p.beta != 0.0) {
#pragma unroll
                        for (int r = 0; r < R; r++)
                            if (row0 + r < N) y[r] = fma(p.beta, yp[r], y[r]);
                    }
                    if (full && p.vec_ok) {
                        if (R == 4) stg256(yp, y[0], y[1], y[2], y[3]);
                        else if (R == 2) stg_stream2(yp, y[0], y[1]);
                        else stg_stream1(yp, y[0]);
                    } else {
#pragma unroll
                        for (int r = 0; r < R; r++)
                            if (row0 + r < N) stg_stream1(yp + r, y[r]);
                    }
                }
            }
        }
        if (SMEM && DB) {
            cp_async_wait_all();  // the next tile has landed ...
            __syncthreads();      // ... and every warp is done with the current buffer
            buf ^= 1;
        }
    }
}

template <int LOGR, bool SMEM, bool DB>
static int32_t launch_sparse_r(sbp_ctx* ctx, const SparseParams& p) {
    auto kern = k_sparse<LOGR, SMEM, DB>;
    const size_t smem = SMEM ? (size_t)p.n_slots * kT * sizeof(double) * (DB ? 2 : 1) : 0;
    if (SMEM) SBP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // warps per CTA: balance the slices over the warps of a CTA and keep as many warps resident as possible
    int best_w = 8, per_sm = 1;
    double best_score = -1.0;
    for (int w = 2; w <= kSparseMaxThreads / 32; w++) {
        int ctas = 0;
        SBP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, w * 32, smem));
        if (ctas < 1) continue;
        const int rounds = (p.n_slices + w - 1) / w;
        const double eff = (double)p.n_slices / ((double)rounds * w);
        const double score = eff * std::min(ctas * w, 32) + 1e-3 * w;
        if (score > best_score) {
            best_score = score;
            best_w = w;
            per_sm = ctas;
        }
    }
    const int THREADS = best_w * 32;
    const long long ntiles = (p.B + kT - 1) / kT;
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)ctx->sm_count * per_sm));
    kern<<<grid, THREADS, smem, ctx->stream>>>(p);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

int32_t launch_sparse_ex(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                         double beta, const sbp_op* adv, const double* g, int64_t g_stride,
                         const double* colscale, const Sat1D* sat) {
    SparseParams p;
    p.U = U;
    p.dU = dU;
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
    if (sat) p.sat = *sat;
    const int R = 1 << op->logR;
    p.n_slots = ((op->N + 2 * R - 1) / (2 * R)) * (2 * R);
    // vector stores need every lane's R rows aligned to R*8 bytes for every instance
    p.vec_ok = (op->N % R == 0) && ((reinterpret_cast<uintptr_t>(dU) & (R * 8 - 1)) == 0);
    const size_t budget = (size_t)ctx->smem_optin - 2048;
    const bool smem_ok = (size_t)p.n_slots * kT * sizeof(double) <= budget;
    p.slice_ptr = op->d_slice_ptr;
    if (smem_ok) {
        p.cols = op->d_sell_col;  // slot indices
        p.vals = op->d_sell_val;
        // double buffering (tuning knob SBP_SPARSE_DB=0/1).  Measured on config 3 (profiles/r1c_sparse_sweep.txt): two
        // single-buffered CTAs per SM overlap staging and compute as well as one double-buffered CTA does, and keep
        // more warps resident, so it is off by default.
        const int db_env = ctx->tune[T_SPARSE_DB];
        const bool db = db_env && 2 * (size_t)p.n_slots * kT * sizeof(double) <= budget;
        if (db) {
            switch (op->logR) {
                case 0: return launch_sparse_r<0, true, true>(ctx, p);
                case 1: return launch_sparse_r<1, true, true>(ctx, p);
                default: return launch_sparse_r<2, true, true>(ctx, p);
            }
        }
        switch (op->logR) {
            case 0: return launch_sparse_r<0, true, false>(ctx, p);
            case 1: return launch_sparse_r<1, true, false>(ctx, p);
            default: return launch_sparse_r<2, true, false>(ctx, p);
        }
    }
    p.cols = op->d_col;  // same layout, plain node indices
    p.vals = op->d_sell_val;
    switch (op->logR) {
        case 0: return launch_sparse_r<0, false, false>(ctx, p);
        case 1: return launch_sparse_r<1, false, false>(ctx, p);
        default: return launch_sparse_r<2, false, false>(ctx, p);
    }
}

int32_t launch_sparse(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                      double beta, const sbp_op* adv, const double* g, int64_t g_stride, double* quantities) {
    int32_t rc = launch_sparse_ex(ctx, op, dU, U, B, alpha, beta, adv, g, g_stride, nullptr, nullptr);
    if (rc) return rc;
    if (adv && quantities) return launch_adv2d_epilogue(ctx, adv, dU, U, B, g, g_stride, quantities, false);
    return SBP_OK;
}

}  // namespace sbp
