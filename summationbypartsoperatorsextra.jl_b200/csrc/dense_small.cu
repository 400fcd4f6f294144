// dense_small.cu -- K1/K4: batched apply of a small dense operator (N <= 128) or a periodic table of
// such operators, operator fragments resident in shared memory, FP64 tensor-core (DMMA.8x8x4) tiles.
//
// Math per instance (src/polynomialbases_operators.jl:148, src/subcell_operators.jl:279, upstream
// MatrixDerivativeOperator):      y = alpha' * D u + beta * y ,  alpha' = alpha*scale
// Batched as a GEMM with the INSTANCES on the M axis:   Y[inst][i] = sum_k U[inst][k] * D[i][k]
//   A operand (row-major 8x4 fragments) = 8 instances x 4 nodes straight from the batch array
//                                          (instance-contiguous layout == "row-major" A)
//   B operand (col-major 4x8 fragments) = D^T, pre-permuted on the host into fragment order and
//                                          staged once per CTA into shared memory with a TMA bulk copy
//   C fragment: a lane ends up with 2 output nodes of one instance.
// The contraction order and the output-row order are free, so the host picks the fragment "layout" that gives the
// widest aligned global accesses:
//   layout 2 (N % 8 == 0):  k-step ks, lane t <-> node 8(ks>>1)+2t+(ks&1);  n-tile nt, column n <-> row 8nt+n
//                           -> one 128-bit load feeds two k-steps, a lane stores 2 consecutive rows (128-bit)
//   layout 1 (any N):       k-step ks, lane t <-> node 4ks+t;               column n <-> row 8nt+(n>>1)+4(n&1)
//                           -> 64-bit accesses, the 4 lanes of a group cover 32 contiguous bytes per instance
//   (layout 4, N % 16 == 0, 256-bit accesses: dense_small_v4.cu)
// A warp owns work items (tile of 8*MT slots, operator gi); the raw loads of the NEXT item are issued right after the
// current item's values are unpacked, so HBM latency hides behind the DMMA chain.
// BAND: for banded operators (FSBP with `bandwidth`, finite-difference SBP operators stored dense as in the
// reference) every fragment (k-pair kp, n-tile nt) with |kp - nt| > BAND is entirely zero; those DMMAs are removed
// at compile time (a runtime bit-mask predicate was tried first: predicated-off DMMAs still cost their issue and the
// kernel did not speed up).  The dropped products are exact zeros and the accumulation order of the remaining ones
// is unchanged, so the result is bitwise the same as multiplying the stored zeros (for finite inputs).
#include <algorithm>
#include "common.cuh"
#include "ptx.cuh"

namespace sbp {

struct DenseSmallParams {
    const double* __restrict__ U;
    double* __restrict__ dU;
    const double* __restrict__ frag;     // G * KS*NT*32
    const double* __restrict__ wfrag;    // G * KS*4  (weights in A-fragment k order) or nullptr
    const double* __restrict__ cfrag;    // KS*4 column scaling (a .* u of the 1D advection) or nullptr
    int N;
    int G;
    long long B;
    double alpha;
    double beta;
    double* energy;   // per instance or nullptr
    double* partial;  // per-CTA energy partial sums or nullptr
    Sat1D sat;        // fused 1D advection SATs (rows 0 and N-1)
};

enum { MODE_VEC2 = 0, MODE_SCALAR2 = 1, MODE_SCALAR1 = 2 };

// raw A data of one work item: MT m-tiles x NT pairs (k-steps 2kp, 2kp+1) per lane
template <int NT, int MT, int MODE>
__device__ __forceinline__ void load_item(double2 (&raw)[MT][NT], const double* __restrict__ U, long long tile, int gi,
                                          int G, int g, int t, int N, long long B) {
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
        const long long inst = (tile * (8 * MT) + mt * 8 + g) * G + gi;
        const bool valid = inst < B;
        const double* up = U + inst * N;
#pragma unroll
        for (int kp = 0; kp < NT; kp++) {
            if (MODE == MODE_VEC2) {
                raw[mt][kp] = valid ? ldg_stream2(up + 8 * kp + 2 * t) : make_double2(0.0, 0.0);
            } else if (MODE == MODE_SCALAR2) {
                const int k0 = 8 * kp + 2 * t;
                raw[mt][kp].x = (valid && k0 < N) ? ldg_stream1(up + k0) : 0.0;
                raw[mt][kp].y = (valid && k0 + 1 < N) ? ldg_stream1(up + k0 + 1) : 0.0;
            } else {
                const int k0 = 8 * kp + t;  // k-step 2kp -> node 8kp+t, k-step 2kp+1 -> node 8kp+4+t
                raw[mt][kp].x = (valid && k0 < N) ? ldg_stream1(up + k0) : 0.0;
                raw[mt][kp].y = (valid && k0 + 4 < N) ? ldg_stream1(up + k0 + 4) : 0.0;
            }
        }
    }
}

template <int NT, int MT, int THREADS, int MODE, bool ENERGY, int BAND>
__global__ void __launch_bounds__(THREADS) k_dense_small(const DenseSmallParams p) {
    constexpr int KS = 2 * NT;
    constexpr int FRAG = KS * NT * 32;
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(128) double smem[];
    __shared__ uint64_t bar;
    __shared__ double warp_part[WARPS];

    double* s_frag = smem;                          // G*FRAG
    double* s_wfrag = smem + (size_t)p.G * FRAG;    // G*KS*4 (if ENERGY)
    double* s_cfrag = s_wfrag + (ENERGY ? p.G * KS * 4 : 0);                          // KS*4 (if cfrag)
    const bool cscale = p.cfrag != nullptr;

    // ---- stage the operator once per CTA: TMA bulk copies tracked by one mbarrier ----
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t total = (uint32_t)((size_t)p.G * FRAG * sizeof(double));
        mbar_arrive_expect_tx(&bar, total);
        uint32_t off = 0;
        while (off < total) {
            uint32_t n = total - off < 32768u ? total - off : 32768u;
            bulk_g2s(reinterpret_cast<char*>(s_frag) + off, reinterpret_cast<const char*>(p.frag) + off, n, &bar);
            off += n;
        }
    }
    if (ENERGY)
        for (int i = threadIdx.x; i < p.G * KS * 4; i += THREADS) s_wfrag[i] = p.wfrag[i];
    if (cscale)
        for (int i = threadIdx.x; i < KS * 4; i += THREADS) s_cfrag[i] = p.cfrag[i];
    __syncthreads();
    mbar_wait(&bar, 0);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int N = p.N, G = p.G;
    const long long B = p.B;
    const long long slots = (B + G - 1) / G;  // instance b = slot*G + gi
    const long long ntiles = (slots + 8 * MT - 1) / (8 * MT);
    const long long tstride = (long long)gridDim.x * WARPS;
    const double alpha = p.alpha, beta = p.beta;
    double e_acc = 0.0;

    long long tile = (long long)blockIdx.x * WARPS + warp;
    int gi = 0;
    double2 raw[MT][NT];
    if (tile < ntiles) load_item<NT, MT, MODE>(raw, p.U, tile, gi, G, g, t, N, B);

    while (tile < ntiles) {
        // ---- A fragments: unpack (.x -> k-step 2kp, .y -> k-step 2kp+1) ----
        double a[MT][KS];
        long long inst[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            inst[mt] = (tile * (8 * MT) + mt * 8 + g) * G + gi;
#pragma unroll
            for (int kp = 0; kp < NT; kp++) {
                a[mt][2 * kp] = raw[mt][kp].x;
                a[mt][2 * kp + 1] = raw[mt][kp].y;
            }
        }
        const int cur_gi = gi;
        // fused 1D SATs: fetch the boundary data and the two end values of every instance NOW, so that their latency
        // hides behind the DMMA chain (the lines were just requested by the fragment loads: L2 hits)
        double sat_bl[MT], sat_br[MT], sat_u0[MT], sat_uN[MT];
        if (p.sat.enabled) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const bool v = inst[mt] < B;
                sat_bl[mt] = v ? p.sat.bc[inst[mt] * p.sat.bc_stride] : 0.0;
                sat_br[mt] = v ? p.sat.bc[inst[mt] * p.sat.bc_stride + 1] : 0.0;
                sat_u0[mt] = v ? p.U[inst[mt] * N] : 0.0;
                sat_uN[mt] = v ? p.U[inst[mt] * N + N - 1] : 0.0;
            }
        }
        // next work item of this warp; its loads fly while the tensor pipe works on the current one
        long long ntile = tile;
        int ngi = gi + 1;
        if (ngi == G) {
            ngi = 0;
            ntile = tile + tstride;
        }
        if (ntile < ntiles) load_item<NT, MT, MODE>(raw, p.U, ntile, ngi, G, g, t, N, B);

        // ---- fused discrete energy  u' P u  (warp-shuffle reduction over the 4 lanes of a group) ----
        if (ENERGY) {
            const double* wf = s_wfrag + cur_gi * KS * 4;
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                double e = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ks++) e = fma(wf[ks * 4 + t] * a[mt][ks], a[mt][ks], e);
                e += __shfl_xor_sync(0xffffffffu, e, 1);
                e += __shfl_xor_sync(0xffffffffu, e, 2);
                if (t == 0 && inst[mt] < B) {
                    if (p.energy) p.energy[inst[mt]] = e;
                    e_acc += e;
                }
            }
        }
        if (cscale) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
                for (int ks = 0; ks < KS; ks++) a[mt][ks] *= s_cfrag[ks * 4 + t];
        }
        // ---- DMMA: C[mt][nt] += A[mt][ks] * Bfrag[ks][nt] ----
        double c[MT][NT][2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++) c[mt][nt][0] = c[mt][nt][1] = 0.0;
        const double* bf = s_frag + (size_t)cur_gi * FRAG + lane;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                if (BAND == 0 || ((ks >> 1) - nt <= BAND && nt - (ks >> 1) <= BAND)) {  // resolved at compile time
                    const double b = bf[(ks * NT + nt) * 32];
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) dmma884(c[mt][nt][0], c[mt][nt][1], a[mt][ks], b);
                }
            }
        }
        // ---- epilogue: lane owns 2 output nodes per n-tile of instance g of each m-tile ----
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            if (inst[mt] >= B) continue;
            double* yp = p.dU + inst[mt] * N;
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                double y0 = alpha * c[mt][nt][0], y1 = alpha * c[mt][nt][1];
                if (p.sat.enabled && (nt == 0 || nt == NT - 1)) {
                    // rows owned by this lane in this n-tile (see the layouts above)
                    const int r0 = (MODE == MODE_SCALAR1) ? 8 * nt + t : 8 * nt + 2 * t;
                    const int r1 = (MODE == MODE_SCALAR1) ? r0 + 4 : r0 + 1;
                    if (r0 == 0) y0 += sat1d_left_v(p.sat, sat_bl[mt], p.sat.a0 * sat_u0[mt]);
                    if (r0 == N - 1) y0 += sat1d_right_v(p.sat, sat_br[mt], p.sat.aN * sat_uN[mt]);
                    if (r1 == N - 1) y1 += sat1d_right_v(p.sat, sat_br[mt], p.sat.aN * sat_uN[mt]);
                }
                if (MODE == MODE_VEC2) {
                    double* q = yp + 8 * nt + 2 * t;
                    if (beta != 0.0) {
                        const double2 o = *reinterpret_cast<const double2*>(q);
                        y0 = fma(beta, o.x, y0);
                        y1 = fma(beta, o.y, y1);
                    }
                    stg_stream2(q, y0, y1);
                } else {
                    const int r0 = (MODE == MODE_SCALAR2) ? 8 * nt + 2 * t : 8 * nt + t;
                    const int r1 = (MODE == MODE_SCALAR2) ? r0 + 1 : r0 + 4;
                    if (r0 < N) {
                        if (beta != 0.0) y0 = fma(beta, yp[r0], y0);
                        stg_stream1(yp + r0, y0);
                    }
                    if (r1 < N) {
                        if (beta != 0.0) y1 = fma(beta, yp[r1], y1);
                        stg_stream1(yp + r1, y1);
                    }
                }
            }
        }
        tile = ntile;
        gi = ngi;
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

template <int NT, int MT, int THREADS, int MODE, bool ENERGY, int BAND>
static int32_t launch_variant(sbp_ctx* ctx, const DenseSmallParams& p, double* energy_sum) {
    constexpr int KS = 2 * NT;
    auto kern = k_dense_small<NT, MT, THREADS, MODE, ENERGY, BAND>;
    size_t smem = (size_t)p.G * KS * NT * 32 * 8 + (ENERGY ? (size_t)p.G * KS * 4 * 8 : 0) + (p.cfrag ? KS * 4 * 8 : 0);
    SBP_REQUIRE((int64_t)smem <= ctx->smem_optin - 1024, SBP_EUNSUPPORTED,
                "operator table needs %zu bytes of shared memory (> %lld available)", smem,
                (long long)ctx->smem_optin);
    SBP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SBP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const long long slots = (p.B + p.G - 1) / p.G;
    const long long ntiles = (slots + 8 * MT - 1) / (8 * MT);
    long long want = (ntiles + (THREADS / 32) - 1) / (THREADS / 32);
    int grid = (int)std::min<long long>((long long)ctx->sm_count * per_sm, std::max<long long>(want, 1));
    DenseSmallParams q = p;
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

template <int NT>
static int32_t dispatch_nt(sbp_ctx* ctx, const DenseSmallParams& p, int mode, bool energy, int band,
                           double* energy_sum) {
    // m-tiles per warp iteration: keep ~64 accumulator doubles per lane at most
    constexpr int MT = (NT <= 2) ? 4 : (NT <= 8 ? 2 : 1);
    constexpr int TH = 128;
    // band variants exist only where they remove work (NT >= 4) and for the apply without fused energy
    constexpr int B1 = (NT >= 4) ? 1 : 0, B2 = (NT >= 6) ? 2 : 0;
#define SBP_DS(M)                                                                                      \
    do {                                                                                               \
        if (energy) return launch_variant<NT, MT, TH, M, true, 0>(ctx, p, energy_sum);                 \
        if (band == 1 && B1) return launch_variant<NT, MT, TH, M, false, B1>(ctx, p, energy_sum);      \
        if (band == 2 && B2) return launch_variant<NT, MT, TH, M, false, B2>(ctx, p, energy_sum);      \
        return launch_variant<NT, MT, TH, M, false, 0>(ctx, p, energy_sum);                            \
    } while (0)
    if (mode == MODE_VEC2) SBP_DS(MODE_VEC2);
    if (mode == MODE_SCALAR2) SBP_DS(MODE_SCALAR2);
    SBP_DS(MODE_SCALAR1);
#undef SBP_DS
}

int32_t launch_dense_small_ex(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                              double beta, double* energy, double* energy_sum, const double* cfrag,
                              const Sat1D* sat) {
    DenseSmallParams p;
    p.U = U;
    p.dU = dU;
    p.frag = op->d_frag;
    p.wfrag = op->d_wfrag;
    p.cfrag = cfrag;
    p.N = op->N;
    p.G = op->G;
    p.B = B;
    p.alpha = alpha * op->scale;
    p.beta = beta;
    p.energy = energy;
    p.partial = nullptr;
    if (sat) p.sat = *sat;
    const bool want_energy = (energy != nullptr) || (energy_sum != nullptr);
    SBP_REQUIRE(!want_energy || op->d_wfrag, SBP_EINVAL, "energy requested but the operator has no weights");
    const int band = op->band;  // 0 = full operator
    const bool skip = band > 0 && !want_energy;
    // 256-bit variant (N % 16 == 0, 32-byte aligned batches, no column scaling, no zero-fragment skipping);
    // SBP_DS_V4 = -1 disables it, otherwise its low bits select the variant (dense_small_v4.cu)
    const int v4 = ctx->tune[T_DS_V4];
    if (v4 >= 0 && !skip && sat == nullptr && op->d_frag4 && cfrag == nullptr &&
        ((reinterpret_cast<uintptr_t>(U) & 31) == 0) &&
        ((reinterpret_cast<uintptr_t>(dU) & 31) == 0))
        return launch_dense_small_v4(ctx, op, dU, U, B, alpha, beta, energy, energy_sum, v4);
    int mode;
    if (op->layout == 1)
        mode = MODE_SCALAR1;
    else
        mode = (((reinterpret_cast<uintptr_t>(U) | reinterpret_cast<uintptr_t>(dU)) & 15) == 0) ? MODE_VEC2 : MODE_SCALAR2;
    switch (op->NT) {
#define SBP_CASE(n) \
    case n:         \
        return dispatch_nt<n>(ctx, p, mode, want_energy, band, energy_sum);
        SBP_CASE(1)
        SBP_CASE(2)
        SBP_CASE(3)
        SBP_CASE(4)
        SBP_CASE(5)
        SBP_CASE(6)
        SBP_CASE(7)
        SBP_CASE(8)
        SBP_CASE(9)
        SBP_CASE(10)
        SBP_CASE(11)
        SBP_CASE(12)
        SBP_CASE(13)
        SBP_CASE(14)
        SBP_CASE(15)
        SBP_CASE(16)
#undef SBP_CASE
    }
    set_error("dense_small: unsupported N=%d", op->N);
    return SBP_EUNSUPPORTED;
}

int32_t launch_dense_small(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                           double beta, bool /*periodic_table*/, double* energy, double* energy_sum) {
    return launch_dense_small_ex(ctx, op, dU, U, B, alpha, beta, energy, energy_sum, nullptr, nullptr);
}

}  // namespace sbp
