// lowstorage.cu -- device-resident time stepping with low-storage 3S*+ Runge-Kutta schemes with an embedded error
// estimator and FSAL (Ranocha, Dalcin, Parsani, Ketcheson 2021): the family of RDPK3SpFSAL49, the integrator of
// examples/RBF_MFSBP_advection.jl:21-27 (`solve(ode, RDPK3SpFSAL49(); ...)`, adaptive, PID step size control).
//
// The coefficients of RDPK3SpFSAL49 live in OrdinaryDiffEqLowStorageRK (third party, not under /root/reference and not in
// this image), so the scheme arrives as a TABLEAU (sbp_tableau_3sp): the Julia shim reads it from the algorithm's cache
// (`SBPB200.lowstorage_tableau(RDPK3SpFSAL49())`), the Python mirror loads the exported file or takes the arrays.  One step,
// in the form OrdinaryDiffEq's LowStorageRK3SpFSAL perform_step uses (S1 = u, S2 = tmp, S3 = uprev):
//     tmp = uprev;  u = tmp + beta_1 dt k_1;  utilde = bhat_1 dt k_1            (k_1 = f(uprev): FSAL of the previous step)
//     for i = 2..s:  k = f(u, t + c_i dt);  tmp += delta_i u;
//                    u = gamma1_i u + gamma2_i tmp + gamma3_i uprev + beta_i dt k;  utilde += bhat_i dt k
//     k_fsal = f(u, t + dt);  utilde += bhat_fsal dt k_fsal
//     EEst = rms(utilde / (abstol + max(|uprev|, |u|) reltol));  PID controller -> accept / reject, next dt
// Every f is one fused RHS launch (sbp_rhs_advection2d / 1d), every update ONE pass over five vectors (k_3sp_update); the
// error norm is a deterministic two-pass reduction.  The host sees one double per step (the norm), as any adaptive driver must.
#include <algorithm>
#include <cmath>
#include <vector>
#include "common.cuh"
#include "ptx.cuh"

namespace sbp {

// tmp += delta u;  u <- g1 u + g2 tmp + g3 uprev + bdt k;  utilde += ehdt k      (utilde may be nullptr: fixed steps)
__global__ void __launch_bounds__(256) k_3sp_update(double* __restrict__ u, double* __restrict__ tmp,
                                                    const double* __restrict__ uprev, const double* __restrict__ k,
                                                    double* __restrict__ utilde, long long n, double delta, double g1,
                                                    double g2, double g3, double bdt, double ehdt) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ui = u[i], ki = k[i];
        const double ti = fma(delta, ui, tmp[i]);
        tmp[i] = ti;
        u[i] = fma(bdt, ki, fma(g3, uprev[i], fma(g2, ti, g1 * ui)));
        if (utilde) utilde[i] = fma(ehdt, ki, utilde[i]);
    }
}

// first stage and FSAL tail:  mode 0: uprev = tmp = u; u += bdt k; utilde = ehdt k      mode 1: utilde += ehdt k
__global__ void __launch_bounds__(256) k_3sp_first(double* __restrict__ u, double* __restrict__ tmp, double* __restrict__ uprev,
                                                   const double* __restrict__ k, double* __restrict__ utilde, long long n,
                                                   double bdt, double ehdt, int mode) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ki = k[i];
        if (mode == 0) {
            const double ui = u[i];
            uprev[i] = ui;
            tmp[i] = ui;
            u[i] = fma(bdt, ki, ui);
            if (utilde) utilde[i] = ehdt * ki;
        } else {
            utilde[i] = fma(ehdt, ki, utilde[i]);
        }
    }
}

// partial[block] = sum_i (utilde_i / (abstol + max(|uprev_i|, |u_i|) reltol))^2      (OrdinaryDiffEq calculate_residuals)
__global__ void __launch_bounds__(256) k_error_norm(const double* __restrict__ utilde, const double* __restrict__ uprev,
                                                    const double* __restrict__ u, long long n, double abstol, double reltol,
                                                    double* __restrict__ partial) {
    __shared__ double wp[8];
    const long long stride = (long long)gridDim.x * blockDim.x;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double r = utilde[i] / (abstol + fmax(fabs(uprev[i]), fabs(u[i])) * reltol);
        s = fma(r, r, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) wp[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += wp[w];
        partial[blockIdx.x] = t;
    }
}

static int grid_for(const sbp_ctx* ctx, long long n) {
    return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16));
}

}  // namespace sbp

using namespace sbp;

extern "C" {

int32_t sbp_error_norm(sbp_ctx* ctx, const double* utilde, const double* uprev, const double* u, int64_t n, double abstol,
                       double reltol, double* h_rms) {
    SBP_REQUIRE(ctx && utilde && uprev && u && h_rms, SBP_EINVAL, "null argument");
    SBP_REQUIRE(n > 0, SBP_EDIM, "empty vectors");
    SBP_CUDA(cudaSetDevice(ctx->device));
    const int grid = grid_for(ctx, n);
    int32_t rc = ensure_scratch(ctx, (size_t)grid + 1);
    if (rc) return rc;
    k_error_norm<<<grid, 256, 0, ctx->stream>>>(utilde, uprev, u, n, abstol, reltol, ctx->scratch);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    rc = launch_sum_partials(ctx, ctx->scratch, grid, ctx->scratch + grid);
    if (rc) return rc;
    double ss = 0.0;
    SBP_CUDA(cudaMemcpyAsync(&ss, ctx->scratch + grid, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    *h_rms = std::sqrt(ss / (double)n);
    return SBP_OK;
}

int32_t sbp_solve_3sp_fsal(sbp_ctx* ctx, const sbp_op* op, double* U, int64_t B, const sbp_tableau_3sp* tab, double t0,
                           double t1, double dt0, int32_t adaptive, double abstol, double reltol, const double* pid,
                           sbp_bc_fn bc, void* bc_user, const double* tstops, int32_t n_tstops, sbp_stop_fn on_stop,
                           void* stop_user, int64_t max_steps, double* h_stats) {
    SBP_REQUIRE(ctx && op && tab && bc, SBP_EINVAL, "null argument");
    SBP_REQUIRE(op->kind == SBP_OP_ADVECTION2D || op->kind == SBP_OP_ADVECTION1D, SBP_EINVAL, "need an advection bundle");
    SBP_REQUIRE(tab->stages >= 1 && tab->beta && tab->bhat, SBP_EINVAL, "bad tableau");
    SBP_REQUIRE(tab->stages == 1 || (tab->gamma1 && tab->gamma2 && tab->gamma3 && tab->delta && tab->c), SBP_EINVAL,
                "tableau arrays missing");
    SBP_REQUIRE(B >= 0, SBP_EDIM, "negative batch size");
    SBP_REQUIRE(t1 >= t0 && dt0 > 0.0, SBP_EINVAL, "need t1 >= t0 and dt0 > 0");
    SBP_REQUIRE(!adaptive || (abstol >= 0.0 && reltol >= 0.0 && abstol + reltol > 0.0 && pid), SBP_EINVAL, "bad tolerances");
    if (B == 0 || t1 == t0) return SBP_OK;
    SBP_REQUIRE(U, SBP_EINVAL, "null state");
    SBP_CUDA(cudaSetDevice(ctx->device));
    const bool two_d = op->kind == SBP_OP_ADVECTION2D;
    const int N = op->N, s = tab->stages;
    const int row = two_d ? N : 2;
    const long long n = (long long)B * N;
    void* w = nullptr;
    int32_t rc = ensure_aux(ctx, 1, (size_t)5 * n * sizeof(double), &w);
    if (rc) return rc;
    double *tmp = static_cast<double*>(w), *uprev = tmp + n, *kf = uprev + n, *ks = kf + n, *ut = ks + n;
    void* d_row = nullptr;
    rc = ensure_aux(ctx, 0, (size_t)row * sizeof(double) * 2, &d_row);  // two slots: consecutive RHS calls never share one
    if (rc) return rc;
    std::vector<double> h_row[2] = {std::vector<double>(row), std::vector<double>(row)};
    int slot = 0;
    const int grid = grid_for(ctx, n);
    // f(u, t) -> out: boundary data from the host callback, uploaded to the slot the previous RHS does not use
    auto rhs = [&](double* out, const double* u, double t) -> int32_t {
        // the pageable staging vector of this slot was consumed two RHS calls ago: its copy has long completed on the stream
        bc(t, h_row[slot].data(), bc_user);
        double* g = static_cast<double*>(d_row) + (size_t)slot * row;
        SBP_CUDA(cudaMemcpyAsync(g, h_row[slot].data(), (size_t)row * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        slot ^= 1;
        return two_d ? sbp_rhs_advection2d(ctx, op, out, u, B, g, 0, nullptr) : sbp_rhs_advection1d(ctx, op, out, u, B, g, 0, nullptr);
    };
    const int k_ctrl = std::max(1, tab->order_for_controller);
    double err[3] = {1.0, 1.0, 1.0};
    double t = t0, dt = dt0;
    int64_t naccept = 0, nreject = 0;
    int next_stop = 0;
    while (next_stop < n_tstops && tstops[next_stop] <= t0) next_stop++;
    rc = rhs(kf, U, t);  // first FSAL evaluation
    if (rc) return rc;
    const double teps = 1e-12 * std::max(1.0, std::fabs(t1));
    while (t < t1 - teps) {
        SBP_REQUIRE(naccept + nreject < max_steps, SBP_EUNSUPPORTED, "sbp_solve_3sp_fsal: max_steps (%lld) reached at t = %g",
                    (long long)max_steps, t);
        double target = t1;
        bool to_stop = false;
        if (next_stop < n_tstops && tstops[next_stop] < t1) target = tstops[next_stop], to_stop = true;
        double h = dt;
        bool clipped = false;
        if (t + h >= target - teps) h = target - t, clipped = true;
        // ---- one step of size h ----
        double* utp = adaptive ? ut : nullptr;
        k_3sp_first<<<grid, 256, 0, ctx->stream>>>(U, tmp, uprev, kf, utp, n, tab->beta[0] * h, tab->bhat[0] * h, 0);
        ctx->launches++;
        for (int i = 1; i < s; i++) {
            rc = rhs(ks, U, t + tab->c[i - 1] * h);
            if (rc) return rc;
            k_3sp_update<<<grid, 256, 0, ctx->stream>>>(U, tmp, uprev, ks, utp, n, tab->delta[i - 1], tab->gamma1[i - 1],
                                                         tab->gamma2[i - 1], tab->gamma3[i - 1], tab->beta[i] * h,
                                                         tab->bhat[i] * h);
            ctx->launches++;
        }
        SBP_CUDA(cudaGetLastError());
        rc = rhs(ks, U, t + h);  // FSAL: becomes k_1 of the next step when this one is accepted
        if (rc) return rc;
        bool accept = true;
        double q = 1.0;
        if (adaptive) {
            if (tab->bhat_fsal != 0.0) {
                k_3sp_first<<<grid, 256, 0, ctx->stream>>>(U, tmp, uprev, ks, ut, n, 0.0, tab->bhat_fsal * h, 1);
                ctx->launches++;
            }
            double eest = 0.0;
            rc = sbp_error_norm(ctx, ut, uprev, U, n, abstol, reltol, &eest);
            if (rc) return rc;
            // PID controller of OrdinaryDiffEq (stepsize_controller!(integrator, ::PIDController, alg))
            eest = std::max(eest, 2.220446049250313e-16);
            err[0] = 1.0 / eest;
            q = std::pow(err[0], pid[0] / k_ctrl) * std::pow(err[1], pid[1] / k_ctrl) * std::pow(err[2], pid[2] / k_ctrl);
            q = 1.0 + std::atan(q - 1.0);  // default_dt_factor_limiter
            accept = q >= 0.81;            // accept_safety
        }
        if (accept) {
            t = clipped ? target : t + h;
            naccept++;
            std::swap(kf, ks);
            if (adaptive) err[2] = err[1], err[1] = err[0];
            dt = adaptive ? h * q : dt0;
            if (clipped && to_stop) {
                next_stop++;
                if (on_stop) {
                    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
                    on_stop(t, stop_user);
                }
            }
        } else {
            nreject++;
            SBP_CUDA(cudaMemcpyAsync(U, uprev, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            dt = h * q;  // kf still holds f(uprev)
        }
    }
    if (h_stats) h_stats[0] = (double)naccept, h_stats[1] = (double)nreject, h_stats[2] = dt;
    return SBP_OK;
}

}  // extern "C"
