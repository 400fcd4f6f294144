// reduce.cu -- K5: P-weighted reductions (integrate / mass / energy / analysis quantities), mass-matrix
// scaling and the Runge-Kutta stage update.  All are streaming, HBM-bound kernels: coalesced 64/128-bit
// accesses, warp-shuffle reductions, deterministic two-pass batch sums (per-CTA partials + one final CTA).
#include <algorithm>
#include "common.cuh"
#include "ptx.cuh"

namespace sbp {

template <int GS>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = GS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- integrate: out[b] = sum_i w_i f(u_i, v_i) ---------------------------------------------------
// A group of GS lanes (GS = 4..32, power of two) owns one instance at a time; lanes stride over nodes.
template <int GS, int MODE>
__global__ void __launch_bounds__(256) k_integrate(const double* __restrict__ U, const double* __restrict__ V,
                                                   const double* __restrict__ w, int N, int G, long long B,
                                                   double* __restrict__ out, double* __restrict__ partial, int power) {
    __shared__ double warp_part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & (GS - 1);
    const long long groups_per_block = 256 / GS;
    const long long gid = (long long)blockIdx.x * groups_per_block + threadIdx.x / GS;
    const long long gstride = (long long)gridDim.x * groups_per_block;
    double acc_total = 0.0;
    // all groups of a warp must iterate together (shuffles): loop on the warp's first group
    const long long first = gid - (lane / GS);
    for (long long b0 = first; b0 < B; b0 += gstride) {
        const long long b = b0 + lane / GS;
        double s = 0.0;
        if (b < B) {
            const double* u = U + b * N;
            const double* wg = w + (size_t)(G > 1 ? (b % G) : 0) * N;
            for (int i = gl; i < N; i += GS) {
                const double ui = ldg_stream1(u + i);
                double f;
                if (MODE == 0) f = ui;
                else if (MODE == 1) f = ui * ui;
                else if (MODE == 2) f = ui * ldg_stream1(V + b * N + i);
                else if (MODE == 3) f = fabs(ui);
                else {  // u^k, k >= 0 (binary powering: the rounding of Base.power_by_squaring, which x -> x^k calls)
                    f = 1.0;
                    double x = ui;
                    for (int k = power; k > 0; k >>= 1) {
                        if (k & 1) f *= x;
                        x *= x;
                    }
                }
                s = fma(wg[i], f, s);
            }
        }
        s = group_sum<GS>(s);
        if (gl == 0 && b < B) {
            if (out) out[b] = s;
            acc_total += s;
        }
    }
    if (partial) {
        acc_total = warp_sum(acc_total);
        if (lane == 0) warp_part[warp] = acc_total;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < 8; i++) t += warp_part[i];
            partial[blockIdx.x] = t;
        }
    }
}

__global__ void __launch_bounds__(256) k_sum_partials(const double* __restrict__ partial, int n, double* out) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

int32_t launch_sum_partials(sbp_ctx* ctx, const double* partials, int n, double* out) {
    k_sum_partials<<<1, 256, 0, ctx->stream>>>(partials, n, out);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

static int pick_gs(int N) {
    int gs = 4;
    while (gs < 32 && gs < N) gs <<= 1;
    return gs;
}

template <int GS>
static int32_t launch_integrate_gs(sbp_ctx* ctx, const double* d_w, int N, int G, const double* U, const double* V,
                                   long long B, int mode, double* out, double* partial, int grid) {
    switch (mode) {
        case 0: k_integrate<GS, 0><<<grid, 256, 0, ctx->stream>>>(U, V, d_w, N, G, B, out, partial, 0); break;
        case 1: k_integrate<GS, 1><<<grid, 256, 0, ctx->stream>>>(U, V, d_w, N, G, B, out, partial, 0); break;
        case 2: k_integrate<GS, 2><<<grid, 256, 0, ctx->stream>>>(U, V, d_w, N, G, B, out, partial, 0); break;
        case 3: k_integrate<GS, 3><<<grid, 256, 0, ctx->stream>>>(U, V, d_w, N, G, B, out, partial, 0); break;
        default: k_integrate<GS, 4><<<grid, 256, 0, ctx->stream>>>(U, V, d_w, N, G, B, out, partial, mode - 16); break;
    }
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

int32_t launch_integrate(sbp_ctx* ctx, const double* d_w, int N, int G, const double* U, const double* V, int64_t B,
                         int mode, double* out, double* out_sum) {
    SBP_REQUIRE((mode >= 0 && mode <= 3) || (mode >= 16 && mode < 16 + 64), SBP_EINVAL,
                "integrate: mode must be 0 (u), 1 (u^2), 2 (u*v), 3 (|u|) or 16 + k (u^k, 0 <= k < 64)");
    SBP_REQUIRE(mode != 2 || V != nullptr, SBP_EINVAL, "integrate: mode 2 needs V");
    const int gs = pick_gs(N);
    const long long groups_per_block = 256 / gs;
    long long want = (B + groups_per_block - 1) / groups_per_block;
    int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)ctx->sm_count * 8));
    double* partial = nullptr;
    if (out_sum) {
        int32_t rc = ensure_scratch(ctx, grid);
        if (rc) return rc;
        partial = ctx->scratch;
    }
    int32_t rc;
    switch (gs) {
        case 4: rc = launch_integrate_gs<4>(ctx, d_w, N, G, U, V, B, mode, out, partial, grid); break;
        case 8: rc = launch_integrate_gs<8>(ctx, d_w, N, G, U, V, B, mode, out, partial, grid); break;
        case 16: rc = launch_integrate_gs<16>(ctx, d_w, N, G, U, V, B, mode, out, partial, grid); break;
        default: rc = launch_integrate_gs<32>(ctx, d_w, N, G, U, V, B, mode, out, partial, grid); break;
    }
    if (rc) return rc;
    if (out_sum) return launch_sum_partials(ctx, partial, grid, out_sum);
    return SBP_OK;
}

// ---- scale_by_(inverse_)mass_matrix! ---------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scale(double* __restrict__ U, const double* __restrict__ w, int N,
                                               long long total, double factor, int inverse) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int i = (int)(idx % N);
        const double u = U[idx];
        // reference: factor * u[i] * w  resp.  factor * u[i] / w   (left-to-right evaluation)
        U[idx] = inverse ? (factor * u) / w[i] : (factor * u) * w[i];
    }
}

int32_t launch_scale(sbp_ctx* ctx, const double* d_w, int N, double* U, int64_t B, double factor, int inverse) {
    const long long total = (long long)B * N;
    int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16));
    k_scale<<<grid, 256, 0, ctx->stream>>>(U, d_w, N, total, factor, inverse);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

// ---- Y = a X + b Y + c Z + d W ------------------------------------------------------------------------
// Stage updates of the explicit Runge-Kutta schemes in ONE pass (SSPRK53's u3 = a30 u + a32 u2 + b32 h f(u2) reads three
// vectors and writes one).  Y is read only when b != 0 (it may be uninitialised memory otherwise); Z, W may be NULL.
// Y may alias X, Z or W (sbp_rhs_advection*_stage allows Uextra == Unext; DeviceBatch.axpby_/lincomb_ accept self as an
// operand): no __restrict__, and coherent streaming loads (ld.global.cs) -- a non-coherent (.nc) load of memory the same
// kernel writes is undefined.  Every thread reads all of its operands before it writes its element.
template <bool VEC>
__global__ void __launch_bounds__(256) k_lincomb(double* Y, const double* X, const double* Z, const double* W,
                                                 long long n, double a, double b, double c, double d) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        const long long n2 = n / 2;
        for (long long i = i0; i < n2; i += stride) {
            const double2 x = ldg_cs2(X + 2 * i);
            double2 y = make_double2(a * x.x, a * x.y);
            if (b != 0.0) {
                const double2 o = reinterpret_cast<const double2*>(Y)[i];
                y.x = fma(b, o.x, y.x), y.y = fma(b, o.y, y.y);
            }
            if (Z) {
                const double2 z = ldg_cs2(Z + 2 * i);
                y.x = fma(c, z.x, y.x), y.y = fma(c, z.y, y.y);
            }
            if (W) {
                const double2 w = ldg_cs2(W + 2 * i);
                y.x = fma(d, w.x, y.x), y.y = fma(d, w.y, y.y);
            }
            reinterpret_cast<double2*>(Y)[i] = y;
        }
    }
    for (long long i = (VEC ? 2 * (n / 2) : 0) + i0; i < n; i += stride) {
        double y = a * X[i];
        if (b != 0.0) y = fma(b, Y[i], y);
        if (Z) y = fma(c, Z[i], y);
        if (W) y = fma(d, W[i], y);
        Y[i] = y;
    }
}

int32_t launch_lincomb(sbp_ctx* ctx, double* Y, const double* X, const double* Z, const double* W, int64_t n, double a,
                       double b, double c, double d) {
    int grid = (int)std::max<long long>(1, std::min<long long>((n / 2 + 255) / 256, (long long)ctx->sm_count * 16));
    const bool aligned = ((reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(X) |
                           reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(W)) & 15) == 0;
    if (aligned)
        k_lincomb<true><<<grid, 256, 0, ctx->stream>>>(Y, X, Z, W, n, a, b, c, d);
    else
        k_lincomb<false><<<grid, 256, 0, ctx->stream>>>(Y, X, Z, W, n, a, b, c, d);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}
int32_t launch_axpbypcz(sbp_ctx* ctx, double* Y, const double* X, const double* Z, int64_t n, double a, double b,
                        double c) {
    return launch_lincomb(ctx, Y, X, Z, nullptr, n, a, b, c, 0.0);
}

// ---- 2D advection: SAT epilogue (when A was applied by a dense kernel) + analysis quantities ----------
// q_consts layout (per node, 3*N doubles): [w | w*b | tau_sum]  with tau_sum_j = sum of tau_i over the boundary
// entries i located at node j (src/conservation_laws/multidimensional_linear_advection.jl:98-105).
template <int GS>
__global__ void __launch_bounds__(256) k_adv2d_epilogue(double* __restrict__ dU, const double* __restrict__ U,
                                                        long long B, int N, const double* __restrict__ bdiag,
                                                        const int* __restrict__ mode,
                                                        const double* __restrict__ qc, const double* __restrict__ g,
                                                        long long g_stride, double* __restrict__ quantities,
                                                        int apply_sat) {
    const int lane = threadIdx.x & 31;
    const int gl = lane & (GS - 1);
    const long long groups_per_block = 256 / GS;
    const long long gid = (long long)blockIdx.x * groups_per_block + threadIdx.x / GS;
    const long long gstride = (long long)gridDim.x * groups_per_block;
    const long long first = gid - (lane / GS);
    for (long long b0 = first; b0 < B; b0 += gstride) {
        const long long b = b0 + lane / GS;
        double s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
        if (b < B) {
            const double* u = U + b * N;
            double* du = dU + b * N;
            const double* gb = g ? g + b * g_stride : nullptr;
            for (int i = gl; i < N; i += GS) {
                const double ui = u[i];
                double dui = du[i];
                const int m = mode[i];
                // tmp1 of set_bc! (:60-71): inflow -> bc value, outflow -> u, interior -> 0
                const double t1 = (m == 1) ? gb[i] : (m == 2 ? ui : 0.0);
                if (apply_sat && m == 1) {
                    dui += bdiag[i] * (ui - t1);
                    du[i] = dui;
                }
                if (quantities) {
                    const double w = qc[i], wb = qc[N + i], tau = qc[2 * N + i];
                    s1 = fma(w, ui, s1);
                    s2 = fma(w, dui, s2);
                    s3 = fma(wb, t1, s3);
                    s4 = fma(w * ui, ui, s4);
                    s5 = fma(w * dui, ui, s5);
                    s6 = fma(ui * wb, ui - 2.0 * t1, s6);
                    s7 = fma(tau * t1, t1, s7);
                }
            }
        }
        if (quantities) {
            s1 = group_sum<GS>(s1), s2 = group_sum<GS>(s2), s3 = group_sum<GS>(s3), s4 = group_sum<GS>(s4);
            s5 = group_sum<GS>(s5), s6 = group_sum<GS>(s6), s7 = group_sum<GS>(s7);
            if (gl == 0 && b < B) {
                double* q = quantities + 7 * b;
                q[0] = s1;               // mass
                q[1] = s2;               // mass_rate
                q[2] = s2 + s3;          // mass_rate_boundary
                q[3] = 0.5 * s4;         // energy
                q[4] = s5;               // energy_rate
                q[5] = s5 + 0.5 * s7;    // energy_rate_boundary
                q[6] = s5 - 0.5 * s6;    // energy_rate_boundary_dissipation
            }
        }
    }
}

int32_t launch_adv2d_epilogue(sbp_ctx* ctx, const sbp_op* adv, double* dU, const double* U, int64_t B,
                              const double* g, int64_t g_stride, double* quantities, bool apply_sat) {
    const int N = adv->N;
    const int gs = pick_gs(N);
    const long long groups_per_block = 256 / gs;
    long long want = (B + groups_per_block - 1) / groups_per_block;
    int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)ctx->sm_count * 8));
#define SBP_L(GSV)                                                                                            \
    k_adv2d_epilogue<GSV><<<grid, 256, 0, ctx->stream>>>(dU, U, B, N, adv->d_b, adv->d_mode, adv->d_q_consts, g, \
                                                         g_stride, quantities, apply_sat ? 1 : 0)
    switch (gs) {
        case 4: SBP_L(4); break;
        case 8: SBP_L(8); break;
        case 16: SBP_L(16); break;
        default: SBP_L(32); break;
    }
#undef SBP_L
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

// ---- 1D advection: SATs at both ends + analysis quantities ---------------------------------------------
// consts (per node, 3*N doubles): [w | w .* (D*a) | a]
__device__ __forceinline__ double godunov(double ul, double ur, double a) { return a > 0.0 ? a * ul : a * ur; }

template <int GS>
__global__ void __launch_bounds__(256) k_adv1d_epilogue(double* __restrict__ dU, const double* __restrict__ U,
                                                        long long B, int N, const double* __restrict__ qc,
                                                        const double* __restrict__ bc, long long bc_stride,
                                                        double* __restrict__ quantities, int apply_sat) {
    const int lane = threadIdx.x & 31;
    const int gl = lane & (GS - 1);
    const long long groups_per_block = 256 / GS;
    const long long gid = (long long)blockIdx.x * groups_per_block + threadIdx.x / GS;
    const long long gstride = (long long)gridDim.x * groups_per_block;
    const long long first = gid - (lane / GS);
    const double w0 = qc[0], wN = qc[N - 1], a0 = qc[2 * N], aN = qc[3 * N - 1];
    for (long long b0 = first; b0 < B; b0 += gstride) {
        const long long b = b0 + lane / GS;
        double s1 = 0, s2 = 0, s4 = 0, s5 = 0, s6 = 0;
        double fl = 0, fr = 0, u0 = 0, uN = 0, lb = 0, rb = 0;
        if (b < B) {
            const double* u = U + b * N;
            double* du = dU + b * N;
            u0 = u[0], uN = u[N - 1];
            lb = bc[b * bc_stride], rb = bc[b * bc_stride + 1];
            fl = godunov(lb, u0, a0);
            fr = godunov(uN, rb, aN);
            if (apply_sat && gl == 0) {
                // upstream SATs: du[1] += (fnum_L - a[1]u[1]) / w[1];  du[end] -= (fnum_R - a[end]u[end]) / w[end]
                if (N == 1) {
                    du[0] = (du[0] + (fl - a0 * u0) / w0) - (fr - aN * uN) / wN;
                } else {
                    du[0] += (fl - a0 * u0) / w0;
                    du[N - 1] -= (fr - aN * uN) / wN;
                }
            }
        }
        __syncwarp();
        if (quantities) {
            if (b < B) {
                const double* u = U + b * N;
                const double* du = dU + b * N;
                for (int i = gl; i < N; i += GS) {
                    const double ui = u[i], dui = du[i], w = qc[i], wda = qc[N + i];
                    s1 = fma(w, ui, s1);
                    s2 = fma(w, dui, s2);
                    s4 = fma(w * ui, ui, s4);
                    s5 = fma(w * dui, ui, s5);
                    s6 = fma(ui * ui, wda, s6);
                }
            }
            s1 = group_sum<GS>(s1), s2 = group_sum<GS>(s2), s4 = group_sum<GS>(s4), s5 = group_sum<GS>(s5);
            s6 = group_sum<GS>(s6);
            if (gl == 0 && b < B) {
                double* q = quantities + 7 * b;
                double erb = s5 + 0.5 * s6;  // analysis_callback.jl:152
                if (a0 > 0.0) erb += 0.5 * (-a0 * lb * lb + aN * uN * uN);
                if (aN < 0.0) erb += 0.5 * (-a0 * u0 * u0 + aN * rb * rb);
                double erbd = erb;
                if (a0 > 0.0) erbd += 0.5 * a0 * (u0 - lb) * (u0 - lb);
                if (aN < 0.0) erbd += 0.5 * aN * (uN - rb) * (uN - rb);
                q[0] = s1;
                q[1] = s2;
                q[2] = s2 - fl + fr;
                q[3] = 0.5 * s4;
                q[4] = s5;
                q[5] = erb;
                q[6] = erbd;
            }
        }
    }
}

int32_t launch_adv1d_epilogue(sbp_ctx* ctx, const sbp_op* adv, double* dU, const double* U, int64_t B,
                              const double* bc, int64_t bc_stride, double* quantities, bool apply_sat) {
    const int N = adv->N;
    const int gs = pick_gs(N);
    const long long groups_per_block = 256 / gs;
    long long want = (B + groups_per_block - 1) / groups_per_block;
    int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)ctx->sm_count * 8));
#define SBP_L(GSV)                                                                                             \
    k_adv1d_epilogue<GSV><<<grid, 256, 0, ctx->stream>>>(dU, U, B, N, adv->d_q_consts, bc, bc_stride, quantities, \
                                                         apply_sat ? 1 : 0)
    switch (gs) {
        case 4: SBP_L(4); break;
        case 8: SBP_L(8); break;
        case 16: SBP_L(16); break;
        default: SBP_L(32); break;
    }
#undef SBP_L
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

}  // namespace sbp
