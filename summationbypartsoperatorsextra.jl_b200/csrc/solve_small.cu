// solve_small.cu -- latency path of the device-resident time stepping (BASELINE config 1: the single-instance
// solve of examples/RBF_FSBP_advection.jl:9-51): a whole fixed-step SSPRK53 integration in ONE launch.
//
// With a handful of instances every stage of the batched path is a kernel launch of a few microseconds that moves a few
// kilobytes: 56 steps x 5 stages = 280 launches, ~1.4 ms however fast the kernels are.  Here one CTA owns one instance for
// the whole solve: the state and the three Runge-Kutta work vectors live in shared memory, a thread owns a row of the RHS,
// stages are separated by __syncthreads, nothing but the boundary table and the operator (CSR or dense rows, read-only,
// L1/L2 resident) is read from global memory between the initial load and the final store of u.  The arithmetic is the
// reference's: du = -D (a .* u) + SATs (upstream 1D semidiscretization) resp. du = -A u + B (u - tmp1)
// (src/conservation_laws/multidimensional_linear_advection.jl:60-81), stages u_next = ca u + cb u_extra + ch du with the
// SSPRK53 coefficients of sbp_solve_ssprk53.
#include <algorithm>
#include "common.cuh"
#include "ptx.cuh"

namespace sbp {

struct SolveSmallParams {
    int N;
    long long B;
    // inner operator: CSR (rowptr != nullptr) or dense row-major N x N
    const int* __restrict__ rowptr;
    const int* __restrict__ col;
    const double* __restrict__ val;
    const double* __restrict__ Dm;
    double scale;
    // 1D bundle: nodal speeds (nullptr: a == 1), end speeds / weights; 2D bundle: diag(B) and the set_bc! modes
    const double* __restrict__ a;
    Sat1D sat;
    const double* __restrict__ bdiag;
    const int* __restrict__ mode;
    double* U;
    const double* __restrict__ bc_table;
    long long bc_row_stride;
    long long nsteps;
    const double* __restrict__ hsteps;  // device copy of the step sizes
};

__constant__ double kSspSmall[11] = {0.355909775063327, 0.644090224936674, 0.367933791638137, 0.632066208361863,
                                     0.237593836598569, 0.762406163401431, 0.377268915331368, 0.377268915331368,
                                     0.242995220537396, 0.238458932846290, 0.287632146308408};

// out = ca * in + cb * extra + ch * RHS(in)   (extra may be nullptr); TWO_D selects the SAT form
template <bool TWO_D>
__device__ __forceinline__ void stage_small(const SolveSmallParams& p, double* out, const double* in, const double* extra,
                                            const double* __restrict__ bc, double ca, double cb, double ch) {
    const int N = p.N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double acc = 0.0;
        if (p.rowptr) {
            const int j1 = __ldg(p.rowptr + i + 1);
            for (int j = __ldg(p.rowptr + i); j < j1; j++) {
                const int c = __ldg(p.col + j);
                const double v = (!TWO_D && p.a) ? __ldg(p.a + c) * in[c] : in[c];
                acc = fma(__ldg(p.val + j), v, acc);
            }
        } else {
            const double* row = p.Dm + (size_t)i * N;
            for (int c = 0; c < N; c++) {
                const double v = (!TWO_D && p.a) ? __ldg(p.a + c) * in[c] : in[c];
                acc = fma(__ldg(row + c), v, acc);
            }
        }
        double rhs = -(p.scale * acc);
        const double ui = in[i];
        if (TWO_D) {
            if (__ldg(p.mode + i) == 1) rhs += __ldg(p.bdiag + i) * (ui - __ldg(bc + i));  // inflow: b_j (u_j - g_j)
        } else {
            if (i == 0) rhs += sat1d_left_v(p.sat, __ldg(bc), p.sat.a0 * ui);
            if (i == N - 1) rhs += sat1d_right_v(p.sat, __ldg(bc + 1), p.sat.aN * ui);
        }
        const double e = extra ? cb * extra[i] : 0.0;
        out[i] = fma(ch, rhs, fma(ca, ui, e));
    }
    __syncthreads();
}

// Register-resident rows: with N <= blockDim a thread owns ONE row for the whole solve, and the operator never changes, so
// the row's (column, value) pairs are read once and kept in registers (R slots, padded with zero entries).  A stage is then R
// independent shared-memory gathers + FMAs (same order as the CSR loop: identical results) instead of a dependent chain of
// cached global loads per non-zero: 280 stages of the N = 101 example drop from ~1 us to ~0.1 us each.
template <bool TWO_D, int R>
__global__ void __launch_bounds__(256) k_solve_ssprk53_small_reg(const SolveSmallParams p) {
    extern __shared__ double sm[];
    const int N = p.N, i = threadIdx.x;
    double *u = sm, *u1 = sm + N, *u2 = sm + 2 * N, *u3 = sm + 3 * N;
    const double* k = kSspSmall;
    int cc[R];
    double cv[R];
    int cn = 0;
    if (i < N) {
        const int j0 = __ldg(p.rowptr + i);
        cn = __ldg(p.rowptr + i + 1) - j0;
#pragma unroll
        for (int j = 0; j < R; j++) {
            cc[j] = j < cn ? __ldg(p.col + j0 + j) : i;
            cv[j] = j < cn ? __ldg(p.val + j0 + j) : 0.0;
        }
    }
    double bd = 0.0;
    bool inflow = false;
    if (TWO_D && i < N) inflow = __ldg(p.mode + i) == 1, bd = __ldg(p.bdiag + i);
    // boundary data of a step (5 stage rows) and its step size are fetched one step AHEAD into registers: a cached global
    // load on the critical path of every stage (L2 latency, all threads wait at the barrier) would cost more than the stage
    const bool left = !TWO_D && i == 0, right = !TWO_D && i == N - 1 && i < N;
    const bool needs_bc = TWO_D ? inflow : (left || right);
    const long long boff = TWO_D ? i : (right ? 1 : 0);
    const double iw = left ? 1.0 / p.sat.w0 : (right ? 1.0 / p.sat.wN : 0.0);  // (f - a u) / w with the reciprocal weight, as k_bsr3
    auto fetch = [&](long long s, double (&g)[5], double& h) {
        h = __ldg(p.hsteps + s);
        if (needs_bc) {
#pragma unroll
            for (int q = 0; q < 5; q++) g[q] = __ldg(p.bc_table + (5 * s + q) * p.bc_row_stride + boff);
        }
    };
    auto stage = [&](double* out, const double* in, const double* extra, double g, double ca, double cb, double ch) {
        if (i < N) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < R; j++) acc = fma(cv[j], j < cn ? in[cc[j]] : 0.0, acc);
            double rhs = -(p.scale * acc);
            const double ui = in[i];
            if (TWO_D) {
                if (inflow) rhs += bd * (ui - g);
            } else {
                if (left) {
                    const double t0 = p.sat.a0 * ui;
                    rhs += ((p.sat.a0 > 0.0 ? p.sat.a0 * g : t0) - t0) * iw;
                }
                if (right) {
                    const double tN = p.sat.aN * ui;
                    rhs -= ((p.sat.aN > 0.0 ? tN : p.sat.aN * g) - tN) * iw;
                }
            }
            const double e = extra ? cb * extra[i] : 0.0;
            out[i] = fma(ch, rhs, fma(ca, ui, e));
        }
        __syncthreads();
    };
    for (long long b = blockIdx.x; b < p.B; b += gridDim.x) {
        double* Ug = p.U + b * N;
        if (i < N) u[i] = Ug[i];
        double g[5] = {0, 0, 0, 0, 0}, gn[5] = {0, 0, 0, 0, 0}, h = 0.0, hn = 0.0;
        if (p.nsteps > 0) fetch(0, g, h);
        __syncthreads();
        for (long long s = 0; s < p.nsteps; s++) {
            if (s + 1 < p.nsteps) fetch(s + 1, gn, hn);
            stage(u1, u, nullptr, g[0], 1.0, 0.0, k[6] * h);
            stage(u2, u1, nullptr, g[1], 1.0, 0.0, k[7] * h);
            stage(u3, u2, u, g[2], k[1], k[0], k[8] * h);
            stage(u1, u3, u, g[3], k[3], k[2], k[9] * h);
            stage(u, u1, u2, g[4], k[5], k[4], k[10] * h);
#pragma unroll
            for (int q = 0; q < 5; q++) g[q] = gn[q];
            h = hn;
        }
        if (i < N) Ug[i] = u[i];
        __syncthreads();
    }
}

template <bool TWO_D>
__global__ void __launch_bounds__(256) k_solve_ssprk53_small(const SolveSmallParams p) {
    extern __shared__ double sm[];
    const int N = p.N;
    double *u = sm, *u1 = sm + N, *u2 = sm + 2 * N, *u3 = sm + 3 * N;
    const double* k = kSspSmall;
    for (long long b = blockIdx.x; b < p.B; b += gridDim.x) {
        double* Ug = p.U + b * N;
        for (int i = threadIdx.x; i < N; i += blockDim.x) u[i] = Ug[i];
        __syncthreads();
        for (long long s = 0; s < p.nsteps; s++) {
            const double h = __ldg(p.hsteps + s);
            const double* bc = p.bc_table + 5 * s * p.bc_row_stride;
            stage_small<TWO_D>(p, u1, u, nullptr, bc, 1.0, 0.0, k[6] * h);
            stage_small<TWO_D>(p, u2, u1, nullptr, bc + p.bc_row_stride, 1.0, 0.0, k[7] * h);
            stage_small<TWO_D>(p, u3, u2, u, bc + 2 * p.bc_row_stride, k[1], k[0], k[8] * h);
            stage_small<TWO_D>(p, u1, u3, u, bc + 3 * p.bc_row_stride, k[3], k[2], k[9] * h);
            stage_small<TWO_D>(p, u, u1, u2, bc + 4 * p.bc_row_stride, k[5], k[4], k[10] * h);
        }
        for (int i = threadIdx.x; i < N; i += blockDim.x) Ug[i] = u[i];
        __syncthreads();
    }
}

// true when the latency path applies (small batch, state fits in shared memory, operator has a row-wise image)
bool solve_small_applies(const sbp_ctx* ctx, const sbp_op* op, int64_t B) {
    const sbp_op* inner = op->inner;
    if (!inner) return false;
    const bool csr = inner->kind == SBP_OP_SPARSE && inner->d_csr_rowptr != nullptr;
    const bool dense = inner->kind == SBP_OP_DENSE && inner->d_D != nullptr && inner->N <= 512;
    if (!csr && !dense) return false;
    if ((size_t)4 * op->N * sizeof(double) > (size_t)ctx->smem_optin - 1024) return false;
    // crossover against one launch per stage (measured, profiles/r2_solve_small.txt): a CTA per instance wins while the batch
    // does not fill the machine many times over
    const int64_t limit = (int64_t)ctx->sm_count * (op->N <= 256 ? 16 : 2);
    return B <= limit;
}

int32_t launch_solve_small(sbp_ctx* ctx, const sbp_op* op, double* U, int64_t B, const double* bc_table,
                           int64_t bc_row_stride, int64_t nsteps, const double* d_hsteps) {
    const sbp_op* inner = op->inner;
    SolveSmallParams p{};
    p.N = op->N;
    p.B = B;
    if (inner->kind == SBP_OP_SPARSE) {
        p.rowptr = inner->d_csr_rowptr, p.col = inner->d_csr_col, p.val = inner->d_csr_val;
    } else {
        p.Dm = inner->d_D;
    }
    p.scale = inner->scale;
    p.U = U;
    p.bc_table = bc_table, p.bc_row_stride = bc_row_stride, p.nsteps = nsteps, p.hsteps = d_hsteps;
    const bool two_d = op->kind == SBP_OP_ADVECTION2D;
    if (two_d) {
        p.bdiag = op->d_b, p.mode = op->d_mode;
    } else {
        p.a = op->d_a ? op->d_q_consts + 2 * (size_t)op->N : nullptr;  // plain nodal speeds (q_consts = [w | w D a | a])
        p.sat.a0 = op->sat_a0, p.sat.aN = op->sat_aN, p.sat.w0 = op->sat_w0, p.sat.wN = op->sat_wN, p.sat.enabled = 1;
    }
    const int threads = op->N <= 128 ? 128 : 256;  // (N <= 256: one row per thread)
    const size_t smem = (size_t)4 * op->N * sizeof(double);
    const int grid = (int)std::min<int64_t>(B, (int64_t)ctx->sm_count * 16);
    constexpr int R = 16;
    if (p.rowptr && !p.a && op->N <= threads && inner->csr_max_row <= R) {  // a row per thread, kept in registers
        auto kreg = two_d ? k_solve_ssprk53_small_reg<true, R> : k_solve_ssprk53_small_reg<false, R>;
        kreg<<<grid, threads, smem, ctx->stream>>>(p);
        ctx->launches++;
        SBP_CUDA(cudaGetLastError());
        return SBP_OK;
    }
    auto kern = two_d ? k_solve_ssprk53_small<true> : k_solve_ssprk53_small<false>;
    if (smem > 48 * 1024) SBP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, ctx->stream>>>(p);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

}  // namespace sbp
