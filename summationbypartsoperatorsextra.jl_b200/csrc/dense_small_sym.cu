// dense_small_sym.cu -- K1 for CENTRO-ANTISYMMETRIC operators (D[N-1-i][N-1-j] == -D[i][j]): every first-derivative
// SBP operator on a node set that is symmetric about the cell centre has this property (Lobatto/Gauss Legendre of
// src/polynomialbases_operators.jl, the classical finite-difference SBP operators, ...).  The even/odd split
//        s_j = u_j + u_{N-1-j},  d_j = u_j - u_{N-1-j}                (j < h = N/2)
//        p   = E s,  q = O d      with  E_ij = (D_ij + D_i,N-1-j)/2,  O_ij = (D_ij - D_i,N-1-j)/2   (i, j < h)
//        y_i = p_i + q_i,         y_{N-1-i} = q_i - p_i
// needs two h x h products instead of one N x N product: HALF the FP64 work.  At N = 64 that turns the apply from
// DMMA-bound (82% tensor-pipe active, profiles/r1a_config2_dense_small_ncu.txt) into an HBM-bound stream.
// Same DMMA formulation as dense_small.cu (instances on the M axis, operator fragments in smem via TMA bulk copy).
#include <algorithm>
#include "common.cuh"
#include "ptx.cuh"

namespace sbp {

struct DenseSymParams {
    const double* __restrict__ U;
    double* __restrict__ dU;
    const double* __restrict__ frag;   // [E | O] fragments, each KSH*NTH*32 doubles
    const double* __restrict__ wsym;   // weights: [w_lo (h, fragment order) | w_hi (h, fragment order)] or nullptr
    int N;
    long long B;
    double alpha;
    double beta;
    double* energy;
    double* partial;
    Sat1D sat;  // fused 1D advection SATs (unit speed only: a variable speed breaks the symmetry)
};

__device__ __forceinline__ int phys_k_sym(int ks, int t) { return 8 * (ks >> 1) + 2 * t + (ks & 1); }

template <int NTH, int MT>
struct RawTile {
    double2 lo[MT][NTH];  // (u[j0], u[j0+1]),           j0 = 8kp + 2t
    double2 hi[MT][NTH];  // (u[N-2-j0], u[N-1-j0])  = mirrors of (j0+1, j0)
};

template <int NTH, int MT>
__device__ __forceinline__ void load_raw(RawTile<NTH, MT>& r, const double* __restrict__ U, long long tile, int g, int t,
                                         int N, long long B) {
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
        const long long inst = tile * (8 * MT) + mt * 8 + g;
        const bool valid = inst < B;
        const double* up = U + inst * N;
#pragma unroll
        for (int kp = 0; kp < NTH; kp++) {
            const int j0 = 8 * kp + 2 * t;
            r.lo[mt][kp] = valid ? ldg_stream2(up + j0) : make_double2(0.0, 0.0);
            r.hi[mt][kp] = valid ? ldg_stream2(up + N - 2 - j0) : make_double2(0.0, 0.0);
        }
    }
}

template <int NTH, int MT, int THREADS, bool ENERGY, bool PREFETCH>
__global__ void __launch_bounds__(THREADS) k_dense_small_sym(const DenseSymParams p) {
    constexpr int KSH = 2 * NTH;           // k-steps over the half index range
    constexpr int FRAG = KSH * NTH * 32;   // doubles per half-operator
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(128) double smem[];
    __shared__ uint64_t bar;
    __shared__ double warp_part[WARPS];
    double* s_E = smem;
    double* s_O = smem + FRAG;
    double* s_w = smem + 2 * FRAG;  // 2 * KSH*4 (if ENERGY)

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        constexpr uint32_t total = 2u * FRAG * sizeof(double);
        mbar_arrive_expect_tx(&bar, total);
        uint32_t off = 0;
        while (off < total) {
            const uint32_t n = total - off < 32768u ? total - off : 32768u;
            bulk_g2s(reinterpret_cast<char*>(smem) + off, reinterpret_cast<const char*>(p.frag) + off, n, &bar);
            off += n;
        }
    }
    if (ENERGY)
        for (int i = threadIdx.x; i < 2 * KSH * 4; i += THREADS) s_w[i] = p.wsym[i];
    __syncthreads();
    mbar_wait(&bar, 0);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int N = p.N;
    const long long B = p.B;
    const long long ntiles = (B + 8 * MT - 1) / (8 * MT);
    const long long tstride = (long long)gridDim.x * WARPS;
    const double alpha = p.alpha, beta = p.beta;
    double e_acc = 0.0;

    long long tile = (long long)blockIdx.x * WARPS + warp;
    RawTile<NTH, MT> raw;
    if (PREFETCH && tile < ntiles) load_raw<NTH, MT>(raw, p.U, tile, g, t, N, B);

    for (; tile < ntiles; tile += tstride) {
        if (!PREFETCH) load_raw<NTH, MT>(raw, p.U, tile, g, t, N, B);
        // ---- even/odd parts in A-fragment order (k-step 2kp <-> node j0, k-step 2kp+1 <-> node j0+1) ----
        double as[MT][KSH], ad[MT][KSH];
        double e_t[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            e_t[mt] = 0.0;
#pragma unroll
            for (int kp = 0; kp < NTH; kp++) {
                const double2 lo = raw.lo[mt][kp], hi = raw.hi[mt][kp];
                as[mt][2 * kp] = lo.x + hi.y;
                ad[mt][2 * kp] = lo.x - hi.y;
                as[mt][2 * kp + 1] = lo.y + hi.x;
                ad[mt][2 * kp + 1] = lo.y - hi.x;
                if (ENERGY) {
                    const double* wl = s_w + (2 * kp) * 4 + t;          // w[j0], w[j0+1] at ks = 2kp, 2kp+1
                    const double* wh = s_w + KSH * 4 + (2 * kp) * 4 + t;  // w[N-1-j0], w[N-2-j0]
                    e_t[mt] = fma(wl[0] * lo.x, lo.x, e_t[mt]);
                    e_t[mt] = fma(wl[4] * lo.y, lo.y, e_t[mt]);
                    e_t[mt] = fma(wh[0] * hi.y, hi.y, e_t[mt]);
                    e_t[mt] = fma(wh[4] * hi.x, hi.x, e_t[mt]);
                }
            }
        }
        // end values for the fused SATs: lane t == 0 holds u[0] (lo.x) and u[N-1] (hi.y) of its instance
        double u_first[MT], u_last[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            u_first[mt] = raw.lo[mt][0].x;
            u_last[mt] = raw.hi[mt][0].y;
        }
        double sat_bl[MT], sat_br[MT];  // boundary data fetched early: latency hides behind the DMMA chain
        if (p.sat.enabled) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const long long inst = tile * (8 * MT) + mt * 8 + g;
                sat_bl[mt] = inst < B ? p.sat.bc[inst * p.sat.bc_stride] : 0.0;
                sat_br[mt] = inst < B ? p.sat.bc[inst * p.sat.bc_stride + 1] : 0.0;
            }
        }
        // the raw registers are dead now: refill them with the next tile while the tensor pipe works
        const long long next = tile + tstride;
        if (PREFETCH && next < ntiles) load_raw<NTH, MT>(raw, p.U, next, g, t, N, B);

        if (ENERGY) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                double e = e_t[mt];
                e += __shfl_xor_sync(0xffffffffu, e, 1);
                e += __shfl_xor_sync(0xffffffffu, e, 2);
                const long long inst = tile * (8 * MT) + mt * 8 + g;
                if (t == 0 && inst < B) {
                    if (p.energy) p.energy[inst] = e;
                    e_acc += e;
                }
            }
        }
        double cp[MT][NTH][2], cq[MT][NTH][2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < NTH; nt++) cp[mt][nt][0] = cp[mt][nt][1] = cq[mt][nt][0] = cq[mt][nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KSH; ks++) {
#pragma unroll
            for (int nt = 0; nt < NTH; nt++) {
                const double be = s_E[(ks * NTH + nt) * 32 + lane];
                const double bo = s_O[(ks * NTH + nt) * 32 + lane];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    dmma884(cp[mt][nt][0], cp[mt][nt][1], as[mt][ks], be);
                    dmma884(cq[mt][nt][0], cq[mt][nt][1], ad[mt][ks], bo);
                }
            }
        }
        // ---- epilogue: rows i0 = 8nt+2t, i0+1 and their mirrors N-1-i0, N-2-i0 ----
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const long long inst = tile * (8 * MT) + mt * 8 + g;
            if (inst >= B) continue;
            double* yp = p.dU + inst * N;
#pragma unroll
            for (int nt = 0; nt < NTH; nt++) {
                const int i0 = 8 * nt + 2 * t;
                double y0 = alpha * (cp[mt][nt][0] + cq[mt][nt][0]);
                double y1 = alpha * (cp[mt][nt][1] + cq[mt][nt][1]);
                double z0 = alpha * (cq[mt][nt][0] - cp[mt][nt][0]);  // row N-1-i0
                double z1 = alpha * (cq[mt][nt][1] - cp[mt][nt][1]);  // row N-2-i0
                if (p.sat.enabled && nt == 0 && t == 0) {              // i0 == 0: rows 0 and N-1
                    y0 += sat1d_left_v(p.sat, sat_bl[mt], p.sat.a0 * u_first[mt]);
                    z0 += sat1d_right_v(p.sat, sat_br[mt], p.sat.aN * u_last[mt]);
                }
                if (beta != 0.0) {
                    const double2 o = *reinterpret_cast<const double2*>(yp + i0);
                    const double2 m = *reinterpret_cast<const double2*>(yp + N - 2 - i0);
                    y0 = fma(beta, o.x, y0);
                    y1 = fma(beta, o.y, y1);
                    z1 = fma(beta, m.x, z1);
                    z0 = fma(beta, m.y, z0);
                }
                stg_stream2(yp + i0, y0, y1);
                stg_stream2(yp + N - 2 - i0, z1, z0);
            }
        }
    }
    if (ENERGY && p.partial) {
        e_acc = warp_sum(e_acc);
        if (lane == 0) warp_part[warp] = e_acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < WARPS; w++) s += warp_part[w];
            p.partial[blockIdx.x] = s;
        }
    }
}

template <int NTH, int MT, int THREADS, bool ENERGY, bool PREFETCH>
static int32_t launch_sym_variant(sbp_ctx* ctx, const DenseSymParams& p, double* energy_sum) {
    constexpr int KSH = 2 * NTH;
    auto kern = k_dense_small_sym<NTH, MT, THREADS, ENERGY, PREFETCH>;
    const size_t smem = (size_t)2 * KSH * NTH * 32 * 8 + (ENERGY ? (size_t)2 * KSH * 4 * 8 : 0);
    SBP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SBP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const long long ntiles = (p.B + 8 * MT - 1) / (8 * MT);
    const long long want = (ntiles + (THREADS / 32) - 1) / (THREADS / 32);
    const int grid = (int)std::min<long long>((long long)ctx->sm_count * per_sm, std::max<long long>(want, 1));
    DenseSymParams q = p;
    if (ENERGY && energy_sum) {
        int32_t rc = ensure_scratch(ctx, grid);
        if (rc) return rc;
        q.partial = ctx->scratch;
    }
    kern<<<grid, THREADS, smem, ctx->stream>>>(q);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    if (ENERGY && energy_sum) return launch_sum_partials(ctx, ctx->scratch, grid, energy_sum);
    return SBP_OK;
}

template <int NTH>
static int32_t dispatch_sym(sbp_ctx* ctx, const DenseSymParams& p, bool energy, double* energy_sum, int variant) {
    // variant (tuning knob, see SBP_SYM_VARIANT in api.cu): bit0 = prefetch, bit1 = MT 1 instead of 2, bit2 = 256 threads
    constexpr int MTL = (NTH <= 2) ? 4 : 2;
#define SBP_SYM(MTV, TH, PF)                                                            \
    return energy ? launch_sym_variant<NTH, MTV, TH, true, PF>(ctx, p, energy_sum)     \
                  : launch_sym_variant<NTH, MTV, TH, false, PF>(ctx, p, energy_sum)
    switch (variant & 7) {
        case 0: SBP_SYM(MTL, 128, false);
        case 1: SBP_SYM(MTL, 128, true);
        case 2: SBP_SYM(1, 128, false);
        case 3: SBP_SYM(1, 128, true);
        case 4: SBP_SYM(MTL, 256, false);
        case 5: SBP_SYM(MTL, 256, true);
        case 6: SBP_SYM(1, 256, false);
        default: SBP_SYM(1, 256, true);
    }
#undef SBP_SYM
}

int32_t launch_dense_small_sym(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                               double beta, double* energy, double* energy_sum, int variant, const Sat1D* sat) {
    DenseSymParams p;
    if (sat) p.sat = *sat;
    p.U = U;
    p.dU = dU;
    p.frag = op->d_frag_sym;
    p.wsym = op->d_wsym;
    p.N = op->N;
    p.B = B;
    p.alpha = alpha * op->scale;
    p.beta = beta;
    p.energy = energy;
    p.partial = nullptr;
    const bool want_energy = energy || energy_sum;
    SBP_REQUIRE(!want_energy || op->d_wsym, SBP_EINVAL, "energy requested but the operator has no weights");
    switch (op->N / 16) {
        case 1: return dispatch_sym<1>(ctx, p, want_energy, energy_sum, variant);
        case 2: return dispatch_sym<2>(ctx, p, want_energy, energy_sum, variant);
        case 3: return dispatch_sym<3>(ctx, p, want_energy, energy_sum, variant);
        case 4: return dispatch_sym<4>(ctx, p, want_energy, energy_sum, variant);
        case 5: return dispatch_sym<5>(ctx, p, want_energy, energy_sum, variant);
        case 6: return dispatch_sym<6>(ctx, p, want_energy, energy_sum, variant);
        case 7: return dispatch_sym<7>(ctx, p, want_energy, energy_sum, variant);
        case 8: return dispatch_sym<8>(ctx, p, want_energy, energy_sum, variant);
    }
    set_error("dense_small_sym: unsupported N=%d", op->N);
    return SBP_EUNSUPPORTED;
}

}  // namespace sbp
