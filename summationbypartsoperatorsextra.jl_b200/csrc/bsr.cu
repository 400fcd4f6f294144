// bsr.cu -- K2'/K3': batched apply of a structurally sparse operator on the FP64 tensor cores (block-sparse DMMA) with
// the advection SAT terms fused into the epilogue.
//
// Reference math: y = alpha*scale*D u + beta*y for operators the reference stores dense although they are banded /
// neighbourhood-sparse (ext/function_space_operators.jl:13-14, src/utils/sparsity_patterns.jl:48-60), and the RHS
//     du = -A u + B (u - tmp1)   (src/conservation_laws/multidimensional_linear_advection.jl:60-81).
//
// Why tensor cores for a sparse operator: the scalar kernel (sparse.cu) is bound by shared-memory wavefronts (8 B
// gathered per FMA, bank conflicts from the irregular pattern, operator entries through L1).  Here the operator is cut
// into ROW BLOCKS of 8 rows; the union of a block's columns is split into FRAGMENTS of 4 columns (any 4 columns: the
// A operand is gathered).  One DMMA.8x8x4 multiplies a fragment (8 rows x 4 columns, zero padded) with 8 instances:
//     A operand  lane (g,t) = U[instance 8mt+g][column c_t]      one conflict-free LDS.64 from the staged tile
//     B operand  lane (g,t) = D[row 8rb+g][column c_t]            one LDS.64, reused by the MT m-tiles of the warp
//     C          lane (g,t) = rows 8rb+2t, 8rb+2t+1 of instance 8mt+g  -> one 128-bit store
// so a gathered 8-byte value feeds 16 flops instead of 2, and the access pattern of the gather is the same for every
// operator (no bank conflicts for runs of consecutive columns).
//
// Streaming ring + warp specialisation: a tile of TI = 8*MT instances is NOT staged whole (N = 1296 x 64 instances =
// 663 KB).  The batch is a 2-D tensor [instance][node]; it is streamed through shared memory in BOXES of kBsrBX = 36
// consecutive nodes x TI instances, each fetched by ONE TMA tensor copy (cp.async.bulk.tensor.2d; out-of-range rows
// and nodes are zero filled by the hardware).  36 doubles per row = 4 (mod 16): the 8 x 4 A-fragment gather is
// bank-conflict free without swizzling.  Boxes live in a ring of NBOX slots (slot = virtual box index mod NBOX, the
// virtual index keeps counting across the tiles of a CTA), so every node is read from HBM exactly once.  Row blocks
// are processed in CHUNKS of 4; chunk c needs the nodes [lo_c, hi_c).  One elected PRODUCER thread issues, per chunk,
// the boxes that complete [0, hi_c) plus two bulk copies with the chunk's operator fragments, all signalling the
// "full" mbarrier of a pipeline stage; eight CONSUMER warps (two groups of four, alternating chunks) wait on it, run
// the DMMA chain of one row block each and release the stage through its "empty" mbarrier.  No CTA-wide barrier in
// the steady state: the epilogue of one warp overlaps the DMMAs of the others, and up to three chunks are in flight
// while one is being multiplied.
#include <algorithm>
#include <cuda.h>
#include "common.cuh"
#include "ptx.cuh"

namespace sbp {

constexpr int kBsrCH = 4;       // row blocks per chunk
constexpr int kBsrGroups = 2;   // consumer groups, alternating chunks
constexpr int kBsrMaxStages = 4;
// consumer warps = kBsrGroups * kBsrCH * MH: warp = (group, row block of the chunk, m-half); a warp owns MTW m-tiles
// (8 instances each) of the tile's TI = 8 * MTW * MH instances; + 1 producer warp

struct BsrParams {
    const double* __restrict__ U;
    double* __restrict__ dU;
    const double* __restrict__ val;   // [fragment][32]
    const int* __restrict__ pos;      // [fragment][4]  (box - boxlo[chunk]) << 8 | node % kBsrBX
    const int4* __restrict__ rbmeta;  // [nrb] {first fragment relative to the chunk, fragments, first row of quad A, of quad B}
    const int4* __restrict__ chmeta;  // [nch] {first fragment, end fragment, boxlo, boxhi}: boxes [0, boxhi) of the tile are
                                      // resident once the chunk has been staged; no later chunk touches a box below boxlo
    int N, nrb, nch, nbox, nbt, maxf, nst;  // nbox ring slots, nbt boxes per tile, nst pipeline stages
    int rb_first, rb_last;            // row blocks holding row 0 and row N-1 (1D SATs)
    long long B;
    double alpha, beta;
    // fused 2D SAT: du_j += b_j (u_j - g_j) on inflow rows
    const int* __restrict__ rbmask;   // [nrb] bit r set <=> row r of the block is an inflow node
    const double* __restrict__ bdiag;
    const double* __restrict__ g;
    long long g_stride;
    Sat1D sat;                        // fused 1D SATs (rows 0 and N-1)
};

// TMA 2-D tensor copy: box {kBsrBX nodes, TI instances} at (node x, instance y) -> shared memory, completion on mbarrier
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
            "r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

template <int MTW, int MH, int SAT, bool VEC>
__global__ void __launch_bounds__((kBsrGroups* kBsrCH* MH + 1) * 32)
    k_bsr(const BsrParams p, const __grid_constant__ CUtensorMap tmap) {
    constexpr int MT = MTW;
    constexpr int TI = 8 * MTW * MH;
    constexpr int CONS = kBsrGroups * kBsrCH * MH;  // consumer warps
    constexpr int BX = kBsrBX;
    constexpr int BOX = TI * BX;  // doubles per box
    extern __shared__ __align__(128) double smem[];
    __shared__ uint64_t bar_full[kBsrMaxStages], bar_empty[kBsrMaxStages];
    double* s_u = smem;                                                              // nbox * BOX
    double* s_val = s_u + (size_t)p.nbox * BOX;                                      // nst * maxf * 32
    int* s_pos = reinterpret_cast<int*>(s_val + (size_t)p.nst * p.maxf * 32);        // nst * maxf * 4
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.N, nst = p.nst, nbox = p.nbox;
    const long long ntiles = (p.B + TI - 1) / TI;
    const int stepb = p.nbt % nbox;  // ring offset advance per tile (in boxes)

    if (threadIdx.x == 0) {
        for (int s = 0; s < nst; s++) {
            mbar_init(&bar_full[s], 1);              // the producer's arrive.expect_tx
            mbar_init(&bar_empty[s], CONS);          // lane 0 of every consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == CONS) {
        // ------------------------------- producer: one elected thread issues TMA -------------------------------
        if (lane == 0) {
            int st = 0, round = 0, off = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int b0 = (int)(tile * TI);
                int jb = 0;  // boxes of this tile already requested
                for (int c = 0; c < p.nch; c++) {
                    if (round > 0) mbar_wait(&bar_empty[st], (round - 1) & 1);
                    const int4 cm = __ldg(p.chmeta + c);
                    const int jn = cm.w, f0 = cm.x, f1 = cm.y;
                    const unsigned vbytes = (unsigned)(f1 - f0) * 256u, pbytes = (unsigned)(f1 - f0) * 16u;
                    mbar_arrive_expect_tx(&bar_full[st], (unsigned)(jn - jb) * (unsigned)(BOX * 8) + vbytes + pbytes);
                    for (int j = jb; j < jn; j++) {
                        int slot = off + j % nbox;
                        if (slot >= nbox) slot -= nbox;
                        tma_load_2d(s_u + (size_t)slot * BOX, &tmap, j * BX, b0, &bar_full[st]);
                    }
                    if (f1 > f0) {
                        bulk_g2s(s_val + (size_t)st * p.maxf * 32, p.val + (size_t)f0 * 32, vbytes, &bar_full[st]);
                        bulk_g2s(s_pos + (size_t)st * p.maxf * 4, p.pos + (size_t)f0 * 4, pbytes, &bar_full[st]);
                    }
                    jb = jn;
                    if (++st == nst) st = 0, round++;
                }
                off += stepb;
                if (off >= nbox) off -= nbox;
            }
        }
        return;
    }

    // ---------------------------------------- consumers: DMMA ----------------------------------------
    const int g = lane >> 2, t = lane & 3;
    const int grp = warp / (kBsrCH * MH), slotw = (warp / MH) % kBsrCH, mh = warp % MH;
    const int i0 = mh * 8 * MTW;  // first instance (tile row) of this warp
    const double* su = s_u + (i0 + g) * BX;

    // per-item operator constants (independent of the tile): fetched one item ahead so that their L2 latency is hidden
    struct Meta {
        int frel, nf, r0, blo, mask;
        double sb0, sb1;
        bool active;
    };
    auto load_meta = [&](int c) {
        Meta m;
        const int rb = c * kBsrCH + slotw;
        m.active = rb < p.nrb;
        m.frel = m.nf = m.blo = m.mask = 0;
        m.r0 = N;
        m.sb0 = m.sb1 = 0.0;
        if (m.active) {
            const int4 r = __ldg(p.rbmeta + rb);
            m.frel = r.x, m.nf = r.y;
            m.r0 = (t < 2 ? r.z : r.w - 4) + 2 * t;  // lane owns rows r0, r0 + 1 (pair t of the block)
            m.blo = __ldg(p.chmeta + c).z;
            if (SAT == 1) {
                m.mask = (__ldg(p.rbmask + rb) >> (2 * t)) & 3;
                if (m.mask & 1) m.sb0 = __ldg(p.bdiag + m.r0);
                if (m.mask & 2) m.sb1 = __ldg(p.bdiag + m.r0 + 1);
            }
        }
        return m;
    };

    // The CTA's stream of items = (tile, chunk) pairs.  Item i is multiplied by group i % G; EVERY consumer warp passes
    // every item's "full" barrier in order and arrives on its "empty" barrier: a chunk also reads boxes that were
    // requested together with earlier items (of the other group), so their arrival must have been observed too, and
    // a stage may only be refilled once no warp can still be waiting on its previous phase.
    int c = 0, off = 0, st = 0, round = 0, turn = 0;
    long long tile = blockIdx.x;
    Meta m = load_meta(grp % p.nch);
    auto next_item = [&]() {
        if (++c == p.nch) {
            c = 0;
            tile += gridDim.x;
            off += stepb;
            if (off >= nbox) off -= nbox;
        }
        if (++st == nst) st = 0, round++;
        if (++turn == kBsrGroups) turn = 0;
    };
    for (; tile < ntiles; next_item()) {
        if (turn != grp) {
            mbar_wait(&bar_full[st], round & 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[st]);
            continue;
        }
        int cn = c + kBsrGroups;  // chunk of this warp's next item: its operator constants are prefetched now
        while (cn >= p.nch) cn -= p.nch;
        const Meta mn = load_meta(cn);

        const long long b0 = tile * TI;
        const int nb = (p.B - b0) < TI ? (int)(p.B - b0) : TI;
        const int rb = c * kBsrCH + slotw;
        const int r0 = m.r0, blo = m.blo, mask = m.mask, nf = m.nf;
        int cb = blo % nbox + off;  // ring slot of box blo in this tile
        if (cb >= nbox) cb -= nbox;
        // element offset of (box-relative, column) code q inside the ring
        auto ring_off = [&](int q) {
            int sb = cb + (q >> 8);
            if (sb >= nbox) sb -= nbox;
            return sb * BOX + (q & 255);
        };
        double acc[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) acc[mt][0] = acc[mt][1] = 0.0;
        double u0[MT], u1[MT];  // SAT == 1: own-row values; SAT == 2: the finished SAT terms of rows r0, r0 + 1
#pragma unroll
        for (int mt = 0; mt < MT; mt++) u0[mt] = u1[mt] = 0.0;

        mbar_wait(&bar_full[st], round & 1);
        if (m.active) {
            const double* sv = s_val + (size_t)st * p.maxf * 32 + (size_t)m.frel * 32 + lane;
            const int* sp = s_pos + (size_t)st * p.maxf * 4 + m.frel * 4 + t;
            if (nf > 0) {
                // software pipeline: ring offsets two fragments ahead, operands one fragment ahead
                double b_cur = sv[0], a_cur[MT];
                {
                    const int o0 = ring_off(sp[0]);
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) a_cur[mt] = su[mt * 8 * BX + o0];
                }
                int o_nxt = ring_off(sp[nf > 1 ? 4 : 0]);
                for (int f = 1; f < nf; f++) {
                    const double b_nxt = sv[f * 32];
                    double a_nxt[MT];
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) a_nxt[mt] = su[mt * 8 * BX + o_nxt];
                    o_nxt = ring_off(sp[(f + 1 < nf ? f + 1 : f) * 4]);
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) dmma884(acc[mt][0], acc[mt][1], a_cur[mt], b_cur);
                    b_cur = b_nxt;
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) a_cur[mt] = a_nxt[mt];
                }
#pragma unroll
                for (int mt = 0; mt < MT; mt++) dmma884(acc[mt][0], acc[mt][1], a_cur[mt], b_cur);
            }
            if (SAT == 1 && mask) {  // own-row values (the chunk's own rows are inside its window)
                const int q0 = r0 / BX, q1 = (r0 + 1) / BX;
                const int o0 = ring_off(((q0 - blo) << 8) | (r0 - q0 * BX));
                const int o1 = ring_off(((q1 - blo) << 8) | (r0 + 1 - q1 * BX));
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    u0[mt] = su[mt * 8 * BX + o0];
                    u1[mt] = su[mt * 8 * BX + o1];
                }
            }
            if (SAT == 2 && (rb == p.rb_first || rb == p.rb_last)) {
                // 1D SATs at rows 0 and N-1: one division per instance.  Lane l evaluates them for the warp's
                // instances l (and l + 32) in parallel; the owners of the two rows fetch them by shuffle.
                const int qR = (N - 1) / BX;
                const int oL = ring_off((0 - blo) << 8), oR = ring_off(((qR - blo) << 8) | (N - 1 - qR * BX));
                double sl[(8 * MTW + 31) / 32], sr[(8 * MTW + 31) / 32];
#pragma unroll
                for (int k = 0; k < (8 * MTW + 31) / 32; k++) {
                    const int il = lane + 32 * k;  // instance within the warp's share
                    sl[k] = sr[k] = 0.0;
                    if (il < 8 * MTW && i0 + il < nb) {
                        const long long inst = b0 + i0 + il;
                        const double* up = s_u + (size_t)(i0 + il) * BX;
                        if (rb == p.rb_first) sl[k] = sat1d_left(p.sat, inst, p.sat.a0 * up[oL]);
                        if (rb == p.rb_last) sr[k] = sat1d_right(p.sat, inst, p.sat.aN * up[oR]);
                    }
                }
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const int il = mt * 8 + g;
                    const double vl = __shfl_sync(0xffffffffu, sl[il >> 5], il & 31);
                    const double vr = __shfl_sync(0xffffffffu, sr[il >> 5], il & 31);
                    if (r0 == 0) u0[mt] = vl;
                    if (r0 == N - 1) u0[mt] += vr;
                    if (r0 + 1 == N - 1) u1[mt] = vr;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[st]);  // the stage (and the ring slots behind it) may be refilled
        if (m.active) {
            // ---- epilogue: lane owns rows r0, r0+1 of instance i0 + 8mt + g ----
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const int il = i0 + mt * 8 + g;
                if (il >= nb) continue;
                const long long inst = b0 + il;
                double y0 = p.alpha * acc[mt][0], y1 = p.alpha * acc[mt][1];
                if (SAT == 1 && mask) {
                    const double* gp = p.g + inst * p.g_stride + r0;
                    if (mask & 1) y0 += m.sb0 * (u0[mt] - __ldg(gp));
                    if (mask & 2) y1 += m.sb1 * (u1[mt] - __ldg(gp + 1));
                }
                if (SAT == 2) {
                    y0 += u0[mt];
                    y1 += u1[mt];
                }
                double* yp = p.dU + inst * N + r0;
                if (VEC) {  // N even: r0 + 1 < N whenever r0 < N
                    if (r0 < N) {
                        if (p.beta != 0.0) {
                            const double2 o = *reinterpret_cast<const double2*>(yp);
                            y0 = fma(p.beta, o.x, y0);
                            y1 = fma(p.beta, o.y, y1);
                        }
                        stg_stream2(yp, y0, y1);
                    }
                } else {
                    if (r0 < N) {
                        if (p.beta != 0.0) y0 = fma(p.beta, yp[0], y0);
                        stg_stream1(yp, y0);
                    }
                    if (r0 + 1 < N) {
                        if (p.beta != 0.0) y1 = fma(p.beta, yp[1], y1);
                        stg_stream1(yp + 1, y1);
                    }
                }
            }
        }
        m = mn;
    }
}

static size_t bsr_smem(const sbp_op* op, int MT, int D) {
    return (size_t)op->bsr_nbox[D] * 8 * MT * kBsrBX * 8 + (size_t)(D + 1) * op->bsr_maxf * (32 * 8 + 4 * 4);
}
static bool bsr_fits(const sbp_ctx* ctx, const sbp_op* op, int MT, int D) {
    return bsr_smem(op, MT, D) + 1024 <= (size_t)ctx->smem_optin;
}

// true when the block-sparse tensor-core kernel can run this operator: one m-tile fits in shared memory and the batch
// rows can be described by a TMA tensor map (row pitch N*8 bytes must be a multiple of 16)
bool bsr_available(const sbp_ctx* ctx, const sbp_op* op) {
    return op->d_bsr_val != nullptr && op->N % 2 == 0 && bsr_fits(ctx, op, 1, 0);
}
bool bsr_can_run(const sbp_op* op, const double* U) {
    return op->variant == SBP_KERNEL_BSR_DMMA && (reinterpret_cast<uintptr_t>(U) & 15) == 0;
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(sym);
    }
    return fn;
}

template <int MTW, int MH, int SAT, bool VEC>
static int32_t launch_bsr_k(sbp_ctx* ctx, const BsrParams& p, size_t smem) {
    auto kern = k_bsr<MTW, MH, SAT, VEC>;
    constexpr int TI = 8 * MTW * MH;
    constexpr int THREADS = (kBsrGroups * kBsrCH * MH + 1) * 32;
    // the batch as a 2-D tensor [B][N] of doubles; one box = kBsrBX nodes x TI instances
    PFN_tmapEncodeTiled enc = tmap_encoder();
    SBP_REQUIRE(enc != nullptr, SBP_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)p.N, (cuuint64_t)p.B};
    const cuuint64_t strides[1] = {(cuuint64_t)p.N * 8};
    const cuuint32_t box[2] = {(cuuint32_t)kBsrBX, (cuuint32_t)TI};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(p.U), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SBP_REQUIRE(r == CUDA_SUCCESS, SBP_ECUDA, "cuTensorMapEncodeTiled failed (%d) for N=%d B=%lld", (int)r, p.N, p.B);
    SBP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SBP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const long long ntiles = (p.B + TI - 1) / TI;
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)ctx->sm_count * per_sm));
    kern<<<grid, THREADS, smem, ctx->stream>>>(p, tmap);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

template <int MTW, int MH>
static int32_t launch_bsr_mt(sbp_ctx* ctx, const BsrParams& p, size_t smem, int sat, bool vec) {
    if (sat == 1)
        return vec ? launch_bsr_k<MTW, MH, 1, true>(ctx, p, smem) : launch_bsr_k<MTW, MH, 1, false>(ctx, p, smem);
    if (sat == 2)
        return vec ? launch_bsr_k<MTW, MH, 2, true>(ctx, p, smem) : launch_bsr_k<MTW, MH, 2, false>(ctx, p, smem);
    return vec ? launch_bsr_k<MTW, MH, 0, true>(ctx, p, smem) : launch_bsr_k<MTW, MH, 0, false>(ctx, p, smem);
}

int32_t launch_bsr(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha, double beta,
                   const sbp_op* adv, const double* g, int64_t g_stride, const Sat1D* sat) {
    SBP_REQUIRE(B < (1ll << 31), SBP_EUNSUPPORTED, "block-sparse kernel: batch size must be < 2^31");
    BsrParams p;
    p.U = U;
    p.dU = dU;
    p.val = op->d_bsr_val;
    p.pos = op->d_bsr_pos;
    p.rbmeta = reinterpret_cast<const int4*>(op->d_bsr_rbmeta);
    p.chmeta = reinterpret_cast<const int4*>(op->d_bsr_chmeta);
    p.N = op->N;
    p.nrb = op->bsr_nrb;
    p.nch = (p.nrb + kBsrCH - 1) / kBsrCH;
    p.rb_first = op->bsr_rb_first;
    p.rb_last = op->bsr_rb_last;
    p.nbt = (op->N + kBsrBX - 1) / kBsrBX;
    p.maxf = op->bsr_maxf;
    p.B = B;
    p.alpha = alpha * op->scale;
    p.beta = beta;
    p.rbmask = adv ? adv->d_rbmask : nullptr;
    p.bdiag = adv ? adv->d_b : nullptr;
    p.g = g;
    p.g_stride = g_stride;
    if (sat) p.sat = *sat;
    const int satmode = (adv && adv->d_rbmask) ? 1 : (sat && sat->enabled ? 2 : 0);
    const bool vec = (reinterpret_cast<uintptr_t>(dU) & 15) == 0;  // N is even
    // (instances per tile, pipeline depth): the largest tile that still allows one chunk of lookahead (operator
    // fragments are re-read from L2 once per tile), then the deepest pipeline that fits; small batches use smaller
    // tiles so that every SM gets work.  SBP_BSR_MT / SBP_BSR_DEPTH override (tuning).
    static int forced_mt = -1, forced_d = -1;
    if (forced_mt < 0) {
        const char* e = getenv("SBP_BSR_MT");
        forced_mt = e ? atoi(e) : 0;
        e = getenv("SBP_BSR_DEPTH");
        forced_d = e ? atoi(e) : 99;
    }
    int MT = 8;
    while (MT > 1 && !bsr_fits(ctx, op, MT, 1)) MT >>= 1;
    while (MT > 1 && (B + 8 * MT - 1) / (8 * MT) < ctx->sm_count) MT >>= 1;
    if ((forced_mt == 1 || forced_mt == 2 || forced_mt == 4 || forced_mt == 8) && bsr_fits(ctx, op, forced_mt, 0))
        MT = forced_mt;
    int D = 0;
    for (int d = 1; d < kBsrMaxStages; d++)
        if (d <= forced_d && bsr_fits(ctx, op, MT, d)) D = d;
    SBP_REQUIRE(bsr_fits(ctx, op, MT, D), SBP_EUNSUPPORTED, "block-sparse kernel needs %zu bytes of shared memory",
                bsr_smem(op, MT, D));
    p.nbox = op->bsr_nbox[D];
    p.nst = D + 1;
    const size_t smem = bsr_smem(op, MT, D);
    // MT m-tiles per tile, split over two warps per row block (16 consumer warps) except for single m-tile tiles;
    // SBP_BSR_MH=1 keeps one warp per row block (8 consumer warps, MT m-tiles each)
    static int forced_mh = -1;
    if (forced_mh < 0) {
        const char* e = getenv("SBP_BSR_MH");
        forced_mh = e ? atoi(e) : 0;
    }
    if (forced_mh == 1) {
        switch (MT) {
            case 8: return launch_bsr_mt<8, 1>(ctx, p, smem, satmode, vec);
            case 4: return launch_bsr_mt<4, 1>(ctx, p, smem, satmode, vec);
            case 2: return launch_bsr_mt<2, 1>(ctx, p, smem, satmode, vec);
            default: return launch_bsr_mt<1, 1>(ctx, p, smem, satmode, vec);
        }
    }
    switch (MT) {
        case 8: return launch_bsr_mt<4, 2>(ctx, p, smem, satmode, vec);
        case 4: return launch_bsr_mt<2, 2>(ctx, p, smem, satmode, vec);
        case 2: return launch_bsr_mt<1, 2>(ctx, p, smem, satmode, vec);
        default: return launch_bsr_mt<1, 1>(ctx, p, smem, satmode, vec);
    }
}

}  // namespace sbp
