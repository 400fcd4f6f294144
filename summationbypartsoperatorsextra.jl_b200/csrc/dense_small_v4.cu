// dense_small_v4.cu -- K1/K4 with 256-bit global accesses, for N % 16 == 0 and 32-byte aligned batches.
// Same DMMA formulation as dense_small.cu (instances on the M axis, operator fragments in shared memory), but
//   * a lane loads 4 consecutive nodes (one 32-byte sector) per LDG.E.256: the 4 lanes of a fragment group cover a FULL
//     128-byte line of one instance, so a warp request touches 8 whole lines instead of 8 half lines -> half the L1/LSU
//     wavefronts per byte (what limited the N = 16 operator-table case, profiles/r1b_config4_ncu.txt);
//   * the contraction order and the output-row order are both permuted on the host (fragment image "layout 4"):
//       k-step ks, lane t   <->  node  16*(ks>>2) + 4t + (ks&3)
//       n-tile nt, column n <->  row   16*(nt>>1) + 4*(n>>1) + 2*(nt&1) + (n&1)
//     so that a lane's accumulators of an n-tile PAIR are 4 consecutive output rows -> one STG.E.256.
#include <algorithm>
#include "common.cuh"
#include "ptx.cuh"

namespace sbp {

struct DenseV4Params {
    const double* __restrict__ U;
    double* __restrict__ dU;
    const double* __restrict__ frag;   // G * KS*NT*32 (layout 4)
    const double* __restrict__ wfrag;  // G * KS*4 (layout 4) or nullptr
    int N;
    int G;
    long long B;
    double alpha;
    double beta;
    double* energy;
    double* partial;
};

template <int NT, int MT>
__device__ __forceinline__ void load_item_v4(double4v (&raw)[MT][NT / 2], const double* __restrict__ U, long long tile,
                                             int gi, int G, int g, int t, int N, long long B) {
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
        const long long inst = (tile * (8 * MT) + mt * 8 + g) * G + gi;
        const bool valid = inst < B;
        const double* up = U + inst * N + 4 * t;
#pragma unroll
        for (int kq = 0; kq < NT / 2; kq++) {
            if (valid) {
                raw[mt][kq] = ldg_stream4(up + 16 * kq);
            } else {
                raw[mt][kq].x = raw[mt][kq].y = raw[mt][kq].z = raw[mt][kq].w = 0.0;
            }
        }
    }
}

template <int NT, int MT, int THREADS, bool ENERGY, bool PREFETCH>
__global__ void __launch_bounds__(THREADS) k_dense_small_v4(const DenseV4Params p) {
    static_assert(NT % 2 == 0, "layout 4 needs N % 16 == 0");
    constexpr int KS = 2 * NT;
    constexpr int FRAG = KS * NT * 32;
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(128) double smem[];
    __shared__ uint64_t bar;
    __shared__ double warp_part[WARPS];
    double* s_frag = smem;
    double* s_wfrag = smem + (size_t)p.G * FRAG;

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
            const uint32_t n = total - off < 32768u ? total - off : 32768u;
            bulk_g2s(reinterpret_cast<char*>(s_frag) + off, reinterpret_cast<const char*>(p.frag) + off, n, &bar);
            off += n;
        }
    }
    if (ENERGY)
        for (int i = threadIdx.x; i < p.G * KS * 4; i += THREADS) s_wfrag[i] = p.wfrag[i];
    __syncthreads();
    mbar_wait(&bar, 0);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int N = p.N, G = p.G;
    const long long B = p.B;
    const long long slots = (B + G - 1) / G;
    const long long ntiles = (slots + 8 * MT - 1) / (8 * MT);
    const long long tstride = (long long)gridDim.x * WARPS;
    const double alpha = p.alpha, beta = p.beta;
    double e_acc = 0.0;

    long long tile = (long long)blockIdx.x * WARPS + warp;
    int gi = 0;
    double4v raw[MT][NT / 2];
    if (PREFETCH && tile < ntiles) load_item_v4<NT, MT>(raw, p.U, tile, gi, G, g, t, N, B);

    while (tile < ntiles) {
        if (!PREFETCH) load_item_v4<NT, MT>(raw, p.U, tile, gi, G, g, t, N, B);
        double a[MT][KS];
        long long inst[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            inst[mt] = (tile * (8 * MT) + mt * 8 + g) * G + gi;
#pragma unroll
            for (int kq = 0; kq < NT / 2; kq++) {
                a[mt][4 * kq] = raw[mt][kq].x;
                a[mt][4 * kq + 1] = raw[mt][kq].y;
                a[mt][4 * kq + 2] = raw[mt][kq].z;
                a[mt][4 * kq + 3] = raw[mt][kq].w;
            }
        }
        const int cur_gi = gi;
        long long ntile = tile;
        int ngi = gi + 1;
        if (ngi == G) {
            ngi = 0;
            ntile = tile + tstride;
        }
        if (PREFETCH && ntile < ntiles) load_item_v4<NT, MT>(raw, p.U, ntile, ngi, G, g, t, N, B);

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
                const double b = bf[(ks * NT + nt) * 32];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) dmma884(c[mt][nt][0], c[mt][nt][1], a[mt][ks], b);
            }
        }
        // epilogue: n-tile pair (2m, 2m+1) of lane t = rows 16m + 4t .. 16m + 4t + 3 of instance g
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            if (inst[mt] >= B) continue;
            double* yp = p.dU + inst[mt] * N + 4 * t;
#pragma unroll
            for (int m = 0; m < NT / 2; m++) {
                double y0 = alpha * c[mt][2 * m][0], y1 = alpha * c[mt][2 * m][1];
                double y2 = alpha * c[mt][2 * m + 1][0], y3 = alpha * c[mt][2 * m + 1][1];
                if (beta != 0.0) {
                    const double2 o0 = *reinterpret_cast<const double2*>(yp + 16 * m);
                    const double2 o1 = *reinterpret_cast<const double2*>(yp + 16 * m + 2);
                    y0 = fma(beta, o0.x, y0);
                    y1 = fma(beta, o0.y, y1);
                    y2 = fma(beta, o1.x, y2);
                    y3 = fma(beta, o1.y, y3);
                }
                stg_stream4(yp + 16 * m, y0, y1, y2, y3);
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

// ---- operator-table variant: the MT m-tiles of a work item are MT DIFFERENT operators over the SAME 8 slots ----------
// Instances b = slot*G + gi: an item (8 slots x MT operators) is one contiguous span of 8*G instances when MT == G, and
// its loads are issued back to back, so DRAM sees dense bursts instead of 128-byte lines 1 KB apart
// (the strided scheme ran at 4.4 TB/s where the contiguous one reaches 5.7 TB/s, profiles/r1d_table_sweep.txt).
template <int NT, int MT>
__device__ __forceinline__ void load_item_tab(double4v (&raw)[MT][NT / 2], const double* __restrict__ U, long long tile,
                                              int og, int G, int g, int t, int N, long long B) {
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
        const int gi = og * MT + mt;
        const long long inst = (tile * 8 + g) * G + gi;
        const bool valid = gi < G && inst < B;
        const double* up = U + inst * N + 4 * t;
#pragma unroll
        for (int kq = 0; kq < NT / 2; kq++) {
            if (valid) {
                raw[mt][kq] = ldg_stream4(up + 16 * kq);
            } else {
                raw[mt][kq].x = raw[mt][kq].y = raw[mt][kq].z = raw[mt][kq].w = 0.0;
            }
        }
    }
}

template <int NT, int MT, int THREADS, bool ENERGY, bool PREFETCH>
__global__ void __launch_bounds__(THREADS) k_table_v4(const DenseV4Params p) {
    static_assert(NT % 2 == 0, "layout 4 needs N % 16 == 0");
    constexpr int KS = 2 * NT;
    constexpr int FRAG = KS * NT * 32;
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(128) double smem[];
    __shared__ uint64_t bar;
    __shared__ double warp_part[WARPS];
    double* s_frag = smem;
    double* s_wfrag = smem + (size_t)p.G * FRAG;

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
            const uint32_t n = total - off < 32768u ? total - off : 32768u;
            bulk_g2s(reinterpret_cast<char*>(s_frag) + off, reinterpret_cast<const char*>(p.frag) + off, n, &bar);
            off += n;
        }
    }
    if (ENERGY)
        for (int i = threadIdx.x; i < p.G * KS * 4; i += THREADS) s_wfrag[i] = p.wfrag[i];
    __syncthreads();
    mbar_wait(&bar, 0);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int N = p.N, G = p.G;
    const long long B = p.B;
    const long long slots = (B + G - 1) / G;
    const long long ntiles = (slots + 7) / 8;
    const int ngroups = (G + MT - 1) / MT;
    const long long tstride = (long long)gridDim.x * WARPS;
    const double alpha = p.alpha, beta = p.beta;
    double e_acc = 0.0;

    long long tile = (long long)blockIdx.x * WARPS + warp;
    int og = 0;
    double4v raw[MT][NT / 2];
    if (PREFETCH && tile < ntiles) load_item_tab<NT, MT>(raw, p.U, tile, og, G, g, t, N, B);

    while (tile < ntiles) {
        if (!PREFETCH) load_item_tab<NT, MT>(raw, p.U, tile, og, G, g, t, N, B);
        double a[MT][KS];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int kq = 0; kq < NT / 2; kq++) {
                a[mt][4 * kq] = raw[mt][kq].x;
                a[mt][4 * kq + 1] = raw[mt][kq].y;
                a[mt][4 * kq + 2] = raw[mt][kq].z;
                a[mt][4 * kq + 3] = raw[mt][kq].w;
            }
        const long long cur_tile = tile;
        const int cur_og = og;
        long long ntile = tile;
        int nog = og + 1;
        if (nog == ngroups) {
            nog = 0;
            ntile = tile + tstride;
        }
        if (PREFETCH && ntile < ntiles) load_item_tab<NT, MT>(raw, p.U, ntile, nog, G, g, t, N, B);

        if (ENERGY) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const int gi = cur_og * MT + mt;
                const double* wf = s_wfrag + (gi < G ? gi : 0) * KS * 4;
                double e = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ks++) e = fma(wf[ks * 4 + t] * a[mt][ks], a[mt][ks], e);
                e += __shfl_xor_sync(0xffffffffu, e, 1);
                e += __shfl_xor_sync(0xffffffffu, e, 2);
                const long long inst = (cur_tile * 8 + g) * G + gi;
                if (t == 0 && gi < G && inst < B) {
                    if (p.energy) p.energy[inst] = e;
                    e_acc += e;
                }
            }
        }
        double c[MT][NT][2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++) c[mt][nt][0] = c[mt][nt][1] = 0.0;
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const int gi = cur_og * MT + mt;
            const double* bf = s_frag + (size_t)(gi < G ? gi : 0) * FRAG + lane;
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
#pragma unroll
                for (int nt = 0; nt < NT; nt++)
                    dmma884(c[mt][nt][0], c[mt][nt][1], a[mt][ks], bf[(ks * NT + nt) * 32]);
        }
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const int gi = cur_og * MT + mt;
            const long long inst = (cur_tile * 8 + g) * G + gi;
            if (gi >= G || inst >= B) continue;
            double* yp = p.dU + inst * N + 4 * t;
#pragma unroll
            for (int m = 0; m < NT / 2; m++) {
                double y0 = alpha * c[mt][2 * m][0], y1 = alpha * c[mt][2 * m][1];
                double y2 = alpha * c[mt][2 * m + 1][0], y3 = alpha * c[mt][2 * m + 1][1];
                if (beta != 0.0) {
                    const double2 o0 = *reinterpret_cast<const double2*>(yp + 16 * m);
                    const double2 o1 = *reinterpret_cast<const double2*>(yp + 16 * m + 2);
                    y0 = fma(beta, o0.x, y0);
                    y1 = fma(beta, o0.y, y1);
                    y2 = fma(beta, o1.x, y2);
                    y3 = fma(beta, o1.y, y3);
                }
                stg_stream4(yp + 16 * m, y0, y1, y2, y3);
            }
        }
        tile = ntile;
        og = nog;
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

template <int NT, int MT, int THREADS, bool ENERGY, bool PREFETCH>
static int32_t launch_tab_variant(sbp_ctx* ctx, const DenseV4Params& p, double* energy_sum) {
    constexpr int KS = 2 * NT;
    auto kern = k_table_v4<NT, MT, THREADS, ENERGY, PREFETCH>;
    const size_t smem = (size_t)p.G * KS * NT * 32 * 8 + (ENERGY ? (size_t)p.G * KS * 4 * 8 : 0);
    SBP_REQUIRE((int64_t)smem <= ctx->smem_optin - 1024, SBP_EUNSUPPORTED, "operator table too large for shared memory");
    SBP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SBP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const long long slots = (p.B + p.G - 1) / p.G;
    const long long ntiles = (slots + 7) / 8;
    const long long want = (ntiles + (THREADS / 32) - 1) / (THREADS / 32);
    const int grid = (int)std::min<long long>((long long)ctx->sm_count * per_sm, std::max<long long>(want, 1));
    DenseV4Params q = p;
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

template <int NT, int MT>
static int32_t dispatch_tab(sbp_ctx* ctx, const DenseV4Params& p, bool energy, double* energy_sum, int variant) {
    if (variant & 1)
        return energy ? launch_tab_variant<NT, MT, 128, true, true>(ctx, p, energy_sum)
                      : launch_tab_variant<NT, MT, 128, false, true>(ctx, p, energy_sum);
    return energy ? launch_tab_variant<NT, MT, 128, true, false>(ctx, p, energy_sum)
                  : launch_tab_variant<NT, MT, 128, false, false>(ctx, p, energy_sum);
}

template <int NT, int MT, int THREADS, bool ENERGY, bool PREFETCH>
static int32_t launch_v4_variant(sbp_ctx* ctx, const DenseV4Params& p, double* energy_sum) {
    constexpr int KS = 2 * NT;
    auto kern = k_dense_small_v4<NT, MT, THREADS, ENERGY, PREFETCH>;
    const size_t smem = (size_t)p.G * KS * NT * 32 * 8 + (ENERGY ? (size_t)p.G * KS * 4 * 8 : 0);
    SBP_REQUIRE((int64_t)smem <= ctx->smem_optin - 1024, SBP_EUNSUPPORTED, "operator table too large for shared memory");
    SBP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SBP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const long long slots = (p.B + p.G - 1) / p.G;
    const long long ntiles = (slots + 8 * MT - 1) / (8 * MT);
    const long long want = (ntiles + (THREADS / 32) - 1) / (THREADS / 32);
    const int grid = (int)std::min<long long>((long long)ctx->sm_count * per_sm, std::max<long long>(want, 1));
    DenseV4Params q = p;
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
static int32_t dispatch_v4(sbp_ctx* ctx, const DenseV4Params& p, bool energy, double* energy_sum, int variant) {
    // variant: bit0 prefetch, bit1 half the m-tiles per item (more items in flight per register budget)
    constexpr int MTL = (NT <= 2) ? 4 : (NT <= 8 ? 2 : 1);
    constexpr int MTS = (MTL > 1) ? MTL / 2 : 1;
#define SBP_V4(MTV, PF)                                                                \
    return energy ? launch_v4_variant<NT, MTV, 128, true, PF>(ctx, p, energy_sum)     \
                  : launch_v4_variant<NT, MTV, 128, false, PF>(ctx, p, energy_sum)
    switch (variant & 3) {
        case 0: SBP_V4(MTL, false);
        case 1: SBP_V4(MTL, true);
        case 2: SBP_V4(MTS, false);
        default: SBP_V4(MTS, true);
    }
#undef SBP_V4
}

int32_t launch_dense_small_v4(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                              double beta, double* energy, double* energy_sum, int variant) {
    DenseV4Params p;
    p.U = U;
    p.dU = dU;
    p.frag = op->d_frag4;
    p.wfrag = op->d_wfrag4;
    p.N = op->N;
    p.G = op->G;
    p.B = B;
    p.alpha = alpha * op->scale;
    p.beta = beta;
    p.energy = energy;
    p.partial = nullptr;
    const bool want_energy = energy || energy_sum;
    SBP_REQUIRE(!want_energy || op->d_wfrag4, SBP_EINVAL, "energy requested but the operator has no weights");
    if (op->G > 1 && !(variant & 4)) {  // operator-major items (bit2 of the variant falls back to slot-major items)
        const int G = op->G;
        if (op->NT == 2) {
            if (G <= 2) return dispatch_tab<2, 2>(ctx, p, want_energy, energy_sum, variant);
            if (G <= 4 || G == 12) return dispatch_tab<2, 4>(ctx, p, want_energy, energy_sum, variant);
            return dispatch_tab<2, 8>(ctx, p, want_energy, energy_sum, variant);
        }
        if (op->NT == 4) {
            if (G <= 2) return dispatch_tab<4, 2>(ctx, p, want_energy, energy_sum, variant);
            return dispatch_tab<4, 4>(ctx, p, want_energy, energy_sum, variant);
        }
        if (op->NT == 8) return dispatch_tab<8, 2>(ctx, p, want_energy, energy_sum, variant);
    }
    switch (op->NT) {
        case 2: return dispatch_v4<2>(ctx, p, want_energy, energy_sum, variant);
        case 4: return dispatch_v4<4>(ctx, p, want_energy, energy_sum, variant);
        case 6: return dispatch_v4<6>(ctx, p, want_energy, energy_sum, variant);
        case 8: return dispatch_v4<8>(ctx, p, want_energy, energy_sum, variant);
        case 10: return dispatch_v4<10>(ctx, p, want_energy, energy_sum, variant);
        case 12: return dispatch_v4<12>(ctx, p, want_energy, energy_sum, variant);
        case 14: return dispatch_v4<14>(ctx, p, want_energy, energy_sum, variant);
        case 16: return dispatch_v4<16>(ctx, p, want_energy, energy_sum, variant);
    }
    set_error("dense_small_v4: unsupported N=%d", op->N);
    return SBP_EUNSUPPORTED;
}

}  // namespace sbp
