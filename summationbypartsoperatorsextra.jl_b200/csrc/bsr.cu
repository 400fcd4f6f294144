// bsr.cu -- K2'/K3': batched apply of a structurally sparse operator on the FP64 tensor cores (block-sparse DMMA) with
// the advection SAT terms fused into the epilogue.
//
// Reference math: y = alpha*scale*D u + beta*y for operators the reference stores dense although they are banded /
// neighbourhood-sparse (ext/function_space_operators.jl:13-14, src/utils/sparsity_patterns.jl:48-60), and the RHS
//     du = -A u + B (u - tmp1)   (src/conservation_laws/multidimensional_linear_advection.jl:60-81).
//
// Why tensor cores for a sparse operator: the scalar kernel (sparse.cu) is bound by shared-memory wavefronts (8 B
// gathered per FMA, bank conflicts from the irregular pattern, operator entries through L1).  Here the operator is cut
// into ROW BLOCKS of 8 rows (two aligned quads of 4 rows, paired on the host so that the union of their columns is
// small: a 2 x 4 patch of a 2-D cloud); the union of a block's columns is split into FRAGMENTS of 4 columns (any 4
// columns: the A operand is gathered).  One DMMA.8x8x4 multiplies a fragment (8 rows x 4 columns, zero padded) with 8
// instances:
//     A operand  lane (g,t) = U[instance 8mt+g][column c_t]      one LDS.64 from the staged boxes (conflict free when
//                                                                 the 4 columns differ mod 4 -- arranged by the host)
//     B operand  lane (g,t) = D[block row g][column c_t]          one LDS.64, reused by the MT m-tiles of the warp
//     C          lane (g,t) = block rows 2t, 2t+1 of instance 8mt+g = one aligned pair of output rows
// so a gathered 8-byte value feeds 16 flops instead of 2 and the gather pattern is the same for every operator.
//
// Data movement (all asynchronous, all through the TMA):
//  * The batch is a 2-D tensor [instance][node].  A tile of TI = 8*MT instances is NOT staged whole (N = 1296 x 56
//    instances = 580 KB): it streams through shared memory in BOXES of kBsrBX = 20 consecutive nodes x TI instances,
//    each fetched by ONE tensor copy (cp.async.bulk.tensor.2d; out-of-range instances and nodes are zero filled by
//    the hardware).  20 doubles per row = 4 (mod 16): the 8 x 4 A-fragment gather is bank-conflict free without
//    swizzling.  Boxes live in a ring (slot = virtual box index mod nbox; the virtual index keeps counting across the
//    tiles of a CTA), so every node is read from HBM exactly once.
//  * Row blocks are processed in CHUNKS of kBsrCH = 8; chunk c needs the boxes [boxlo_c, boxhi_c).  Two single-thread
//    PRODUCERS (separate warps) wait on the "empty" mbarrier of a pipeline stage and request the boxes that complete
//    [0, boxhi_c) resp. the chunk's operator fragments (two bulk copies from L2), all signalling the stage's "full"
//    mbarrier.  Eight CONSUMER warps wait on it, run the DMMA chain of one row block each (MT accumulator tiles per
//    warp, operands software-pipelined one fragment ahead), add the SAT terms from the ring and release the stage.
//  * Results leave through TMA as well: each consumer warp deposits its C fragments in a private staging buffer and
//    one elected lane stores the block's two quads as {4 rows x TI instances} tensor boxes (cp.async.bulk.tensor
//    store, bulk async-group); the warp moves on to the next chunk at once.  With plain 16-byte stores every warp
//    of the CTA stalled on the store queue at the end of every chunk (2100 of 5150 cycles per chunk, measured with
//    clock64 stamps) while the tensor pipe idled; beta != 0 or an unaligned dU fall back to that path.
//  * No CTA-wide barrier after the prologue.
//  * Per-item constants of a consumer warp come from one 32-byte record per row block, prepared on the host for the
//    launch's pipeline depth (first fragment, fragment count, rows of the two quads, ring slot of the chunk's window,
//    flags of the rows that carry a 1D SAT) and fetched one item ahead; alpha = -1 (every RHS call) is folded into the
//    fragment loads as a sign-bit XOR.
//  * Odd N: tensor rows are instance PAIRS and the image is that of diag(D, D) (api.cu build_bsr): same data path.
// Measured (profiles/r1h_*): BASELINE config 3 (N = 1296 cloud, 13.3 nnz/row, B = 65536, fused 2D SAT) 0.325 ms = 262
// GDOF/s (scalar kernel: 0.80 ms); banded N = 104 apply 0.139 ms = 391 GDOF/s = 96% of the measured HBM peak; config 1
// batched (N = 101, fused 1D SATs) 0.168 ms = 315 GDOF/s.
#include <algorithm>
#include <cuda.h>
#include "common.cuh"
#include "ptx.cuh"
#include "bsr.cuh"

// -DSBP_BSR_TRACE: per-warp cycle accounting of the consumer phases (clock64 stamps; block 0 prints its totals)
#ifdef SBP_BSR_TRACE
#include <cstdio>
#define TR_DECL long long tr_t = clock64(), tr_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; int tr_items = 0;
#define TR(k) { const long long tr_n = clock64(); tr_acc[k] += tr_n - tr_t; tr_t = tr_n; }
#else
#define TR_DECL
#define TR(k)
#endif

namespace sbp {

constexpr int kBsrMaxStages = 4;
// consumer warps = kBsrGroups * kBsrCH * MH: warp = (group, row block of the chunk, m-half); a warp owns MTW m-tiles
// (8 instances each) of the tile's TI = 8 * MTW * MH instances; + 2 producer warps

// OUT: how results leave the SM.  0: 8-byte stores (dU not 16-byte aligned), 1: 16-byte stores, 2: staged in shared
// memory and written by TMA tensor stores (beta == 0): asynchronous, so the output stream overlaps the next chunk's
// tensor work instead of stalling every warp of the CTA on the store queue at the same time.
// FUSED: the Runge-Kutta stage update of sbp_rhs_advection*_stage rides on the RHS pass (separate instantiations, so that the
// plain kernels keep their register allocation)
template <int MTW, int MH, int SAT, int OUT, bool FUSED = false>
__global__ void __launch_bounds__((kBsrGroups* kBsrCH* MH + 2) * 32)
    k_bsr(const BsrParams p, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out) {
    constexpr bool VEC = OUT == 1 || OUT == 2;
    // Odd N: a tensor row is an instance PAIR (2N doubles, a multiple of 16 bytes) and the operator image is that of
    // diag(D, D) (api.cu build_bsr): nothing in the data path changes, only the SAT terms map a row of the pair back to
    // (instance, node) with p.Nh.
    constexpr int MTS = 8 * kBsrBX;  // smem distance between the m-tiles of a lane
    constexpr int PS = 4;            // ring codes per fragment
    constexpr int MT = MTW;
    constexpr int TI = 8 * MTW * MH;
    constexpr int CONS = kBsrGroups * kBsrCH * MH;  // consumer warps
    // PING (-DSBP_BSR_PING=1, off by default): the KW consumer warps that share an SM sub-partition (warp % 4) pass a token
    // through named barriers so that their DMMA chains are mutually exclusive (one warp's scalar prologue/epilogue would
    // hide behind the other's chain).  Measured on B200: the token is never contended (28 cycles of waiting per item) and
    // the chain of a warp takes ~30 cycles per DMMA with or without it -- the chain is bound by its own operand
    // latencies, not by the pipe -- so the exchange only costs (config 3: 0.338 ms with, 0.325 ms without).
    constexpr int KW = CONS / 4;
    constexpr bool PING = SBP_BSR_PING && kBsrGroups == 1 && CONS % 4 == 0 && KW >= 2 && KW <= 3;
    constexpr int BX = kBsrBX;
    constexpr int BOX = TI * BX;  // doubles per box
    extern __shared__ __align__(128) double smem[];
    __shared__ uint64_t bar_full[kBsrMaxStages], bar_empty[kBsrMaxStages];
    double* s_u = smem;                                                              // nbox * BOX
    double* s_val = s_u + (size_t)p.nbox * BOX;                                      // nst * maxf * 32
    int* s_pos = reinterpret_cast<int*>(s_val + (size_t)p.nst * p.maxf * 32);        // nst * maxf * PS
    // output staging, per consumer warp: [quad A | quad B][8 * MTW instances][4 rows]  (128-byte aligned)
    double* s_out = reinterpret_cast<double*>(s_pos + (((size_t)p.nst * p.maxf * PS + 31) & ~(size_t)31));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.N, nst = p.nst, nbox = p.nbox;
    const long long ntiles = (p.B + TI - 1) / TI;
    const int stepb = p.nbt % nbox;  // ring offset advance per tile (in boxes)
    const bool alpha_m1 = p.alpha == -1.0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < nst; s++) {
            mbar_init(&bar_full[s], 2);              // the two producers' arrive.expect_tx
            mbar_init(&bar_empty[s], CONS);          // lane 0 of every consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp >= CONS) {
        // ------------------------------------------- producers -------------------------------------------
        // Two single-thread producers (two warps, so that they run concurrently): warp CONS requests the boxes of the
        // batch, warp CONS + 1 the operator fragments.  Each TMA-class instruction costs a thread ~200 cycles of issue
        // time and every trip through this loop is on the critical path of the pipeline (a stage cannot be refilled
        // faster than one trip), so the work is split and each loop is kept to a few dozen instructions: ring slot and
        // box coordinates advance incrementally, the chunk constants of the next item are fetched before the wait.
        if (lane == 0) {
            const bool boxes = warp == CONS;
            int st = 0, round = 0;
            int slot = 0;   // ring slot of the next box to request
            const uint32_t s_u32 = smem_u32(s_u), s_val32 = smem_u32(s_val), s_pos32 = smem_u32(s_pos);
            const uint32_t full32 = smem_u32(&bar_full[0]);
            int4 cm = __ldg(p.chmeta);
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int b0 = (int)(tile * TI);
                int jb = 0, x = 0;  // boxes of this tile already requested, node coordinate of the next box
                for (int c = 0; c < p.nch; c++) {
                    const int4 cmn = __ldg(p.chmeta + (c + 1 < p.nch ? c + 1 : 0));
                    if (round > 0) mbar_wait(&bar_empty[st], (round - 1) & 1);
                    const uint32_t bar = full32 + 8u * st;
                    if (boxes) {
                        const int jn = cm.w;
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar),
                                     "r"((unsigned)(jn - jb) * (unsigned)(BOX * 8))
                                     : "memory");
                        for (; jb < jn; jb++, x += BX) {
                            const uint32_t dst = s_u32 + (uint32_t)slot * (uint32_t)(BOX * 8);
                            asm volatile(
                                "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, "
                                "{%2, %3}], [%4];\n" ::"r"(dst),
                                "l"(&tmap), "r"(x), "r"(b0), "r"(bar)
                                : "memory");
                            if (++slot == nbox) slot = 0;
                        }
                    } else {
                        const unsigned nfr = (unsigned)(cm.y - cm.x);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar),
                                     "r"(nfr * (256u + 4u * PS))
                                     : "memory");
                        if (nfr) {
                            asm volatile(
                                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                                    "r"(s_val32 + (uint32_t)st * (uint32_t)(p.maxf * 256)),
                                "l"(p.val + (size_t)cm.x * 32), "r"(nfr * 256u), "r"(bar)
                                : "memory");
                            asm volatile(
                                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                                    "r"(s_pos32 + (uint32_t)st * (uint32_t)(p.maxf * 4 * PS)),
                                "l"(p.pos + (size_t)cm.x * PS), "r"(nfr * 4u * PS), "r"(bar)
                                : "memory");
                        }
                    }
                    if (++st == nst) st = 0, round++;
                    cm = cmn;
                }
                // boxes of the tile that no chunk needs are skipped: keep slot == (virtual box index) mod nbox
                slot += (p.nbt - jb) % nbox;
                if (slot >= nbox) slot -= nbox;
            }
        }
        return;
    }

    // ---------------------------------------- consumers: DMMA ----------------------------------------
    const int g = lane >> 2, t = lane & 3;
    const int grp = warp / (kBsrCH * MH), slotw = (warp / MH) % kBsrCH, mh = warp % MH;
    const int i0 = mh * 8 * MTW;  // first instance (tile row) of this warp
    const double* su = s_u + (i0 + g) * BX;  // lane g owns tensor row 8 mt + g of m-tile mt
    // (instance, node) of row r of a tensor row: r itself, or (half, r - half * Nh) for instance pairs
    const int Nh = p.Nh;
    const bool paired = Nh != N;

    // per-item operator constants (independent of the tile): one 32-byte record per row block, prepared on the host for
    // this launch's pipeline depth and fetched one item ahead (two independent 16-byte loads + the SAT coefficients)
    struct Meta {
        int4 rb;       // {first fragment relative to the chunk, fragments, first row of quad A, of quad B}
        int4 x;        // {ring slot of the chunk's lowest box, lowest box of the chunk's window, -, -}
        double2 sb;    // SAT coefficients of the lane's two rows (0: no SAT)
    };
    const int4* const items = p.item + 2 * slotw;
    const double2* const sbs = reinterpret_cast<const double2*>(p.rb_bdiag) + 4 * slotw + t;
    auto load_meta = [&](int c) {
        Meta m;
        m.rb = __ldg(items + c * (2 * kBsrCH));
        m.x = __ldg(items + c * (2 * kBsrCH) + 1);
        m.sb = make_double2(0.0, 0.0);
        if (SAT == 1) m.sb = __ldg(sbs + c * (4 * kBsrCH));
        return m;
    };

    // The CTA's stream of items = (tile, chunk) pairs.  Item i is multiplied by group i % G; EVERY consumer warp passes
    // every item's "full" barrier in order and arrives on its "empty" barrier: a chunk also reads boxes that were
    // requested together with earlier items (of the other group), so their arrival must have been observed too, and
    // a stage may only be refilled once no warp can still be waiting on its previous phase.
    int c = 0, off = 0, st = 0, round = 0, turn = 0;
    int tile = blockIdx.x, b0 = blockIdx.x * TI;  // B < 2^31 (checked by the launcher)
    const int ntl = (int)ntiles, tstep = gridDim.x * TI, Bi = (int)p.B;
    Meta m = load_meta(grp % p.nch);
    // tensor-pipe token of this warp's scheduler: barrier id (1 + KW * scheduler + k) = "warp k may multiply"
    const int pp_me = 1 + KW * (warp & 3) + (warp >> 2);
    const int pp_next = 1 + KW * (warp & 3) + ((warp >> 2) + 1 == KW ? 0 : (warp >> 2) + 1);
    if (PING && (warp >> 2) == KW - 1) named_bar_arrive(pp_next, 64);  // the first turn belongs to warp k = 0
    auto next_item = [&]() {
        if (++c == p.nch) {
            c = 0;
            tile += gridDim.x;
            b0 += tstep;
            off += stepb;
            if (off >= nbox) off -= nbox;
        }
        if (++st == nst) st = 0, round++;
        if (++turn == kBsrGroups) turn = 0;
    };
    // 1D SATs: reciprocal boundary weights and, when the boundary data are shared by the batch, the data themselves
    double sat_iw0 = 0.0, sat_iwN = 0.0, sat_lb = 0.0, sat_rb = 0.0;
    if (SAT == 2) {
        sat_iw0 = 1.0 / p.sat.w0, sat_iwN = 1.0 / p.sat.wN;
        if (p.sat.bc_stride == 0) sat_lb = __ldg(p.sat.bc), sat_rb = __ldg(p.sat.bc + 1);
    }
    TR_DECL
    for (; tile < ntl; next_item()) {
        if (turn != grp) {
            mbar_wait(&bar_full[st], round & 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[st]);
            continue;
        }
        int cn = c + kBsrGroups;  // chunk of this warp's next item: its operator constants are prefetched now
        while (cn >= p.nch) cn -= p.nch;
        const Meta mn = load_meta(cn);
        TR(0)  // next_item + metadata prefetch

        const int nb = (Bi - b0) < TI ? Bi - b0 : TI;
        const int rb = c * kBsrCH + slotw;
        const int4 rbm = m.rb;
        const bool active = rbm.z < N;  // padding records of the last chunk carry rows >= N
        const double2 sb = m.sb;
        const int r0 = (t < 2 ? rbm.z : rbm.w - 4) + 2 * t;  // lane owns rows r0, r0 + 1 (pair t of the block)
        const int blo = m.x.y, nf = rbm.y;
        // (half, node) of the lane's two rows: for instance pairs row r >= Nh is node r - Nh of the odd instance
        const int h0 = (paired && r0 >= Nh) ? 1 : 0, h1 = (paired && r0 + 1 >= Nh) ? 1 : 0;
        const int n0 = r0 - h0 * Nh, n1 = r0 + 1 - h1 * Nh;
        // b_j != 0 tested on the bit pattern (no trip through the FP64 pipe)
        const int mask = (SAT == 1) ? (((__double_as_longlong(sb.x) << 1) != 0 ? 1 : 0) | ((__double_as_longlong(sb.y) << 1) != 0 ? 2 : 0)) : 0;
        int cb = m.x.x + off;  // ring slot of box blo in this tile
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

        // boundary data shared by the batch (g_stride == 0): requested before the wait, consumed in the epilogue
        double gs0 = 0.0, gs1 = 0.0;
        if (SAT == 1 && mask && p.g_stride == 0) {
            if (mask & 1) gs0 = __ldg(p.g + n0);
            if (mask & 2) gs1 = __ldg(p.g + n1);
        }
        // stage update: the lane's entries of `extra` are requested now and consumed after the chain
        // (measured, N = 1296: 0.415 ms per stage with the loads ahead of the chain, 0.485 ms with the loads in the epilogue)
        constexpr bool PREFETCH_EXTRA = FUSED;
        double2 ex[PREFETCH_EXTRA ? MT : 1];
        const bool with_extra = FUSED && p.cb != 0.0 && active && r0 < N;
        if (PREFETCH_EXTRA && with_extra) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const int il = i0 + 8 * mt + g;
                ex[mt] = *reinterpret_cast<const double2*>(p.extra + (long long)(b0 + (il < nb ? il : 0)) * N + r0);
            }
        }
        TR(1)  // tile/row/ring constants, boundary data
        mbar_wait(&bar_full[st], round & 1);
        TR(2)  // wait for the stage
        {
            const double* sv = s_val + (size_t)st * p.maxf * 32 + (size_t)rbm.x * 32 + lane;
            const int* sp = s_pos + (size_t)st * p.maxf * PS + rbm.x * PS + t;
            const int nfe = active ? nf : 0;
            // software pipeline: ring offsets two fragments ahead, operands one fragment ahead
            // alpha == -1 (every RHS call): the sign is folded into the operator fragment as it is loaded (one integer
            // XOR per fragment, issued under the DMMA chain) instead of 2*MT FP64 negations on the epilogue's critical path
            const int sgn = alpha_m1 ? (int)0x80000000u : 0;
            auto ldb = [&](const double* q) {
                const double v = *q;
                return __hiloint2double(__double2hiint(v) ^ sgn, __double2loint(v));
            };
            double b_cur = 0.0, a_cur[MT];
            int o_nxt = 0;
            if (nfe > 0) {
                b_cur = ldb(sv);
                const int o0 = ring_off(sp[0]);
#pragma unroll
                for (int mt = 0; mt < MT; mt++) a_cur[mt] = su[mt * MTS + o0];
                o_nxt = ring_off(sp[nfe > 1 ? PS : 0]);
            }
            TR(3)  // operand pointers, first operands
            if (PING) named_bar_sync(pp_me, 64);  // first operands are in flight; wait for the pipe
            TR(4)  // wait for the pipe token
            if (nfe > 0) {
                for (int f = 1; f < nfe; f++) {
                    const double b_nxt = ldb(sv + f * 32);
                    double a_nxt[MT];
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) a_nxt[mt] = su[mt * MTS + o_nxt];
                    o_nxt = ring_off(sp[(f + 1 < nfe ? f + 1 : f) * PS]);
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) dmma884(acc[mt][0], acc[mt][1], a_cur[mt], b_cur);
                    b_cur = b_nxt;
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) a_cur[mt] = a_nxt[mt];
                }
#pragma unroll
                for (int mt = 0; mt < MT; mt++) dmma884(acc[mt][0], acc[mt][1], a_cur[mt], b_cur);
            }
            TR(5)  // DMMA chain
            if (PING) named_bar_arrive(pp_next, 64);  // chain issued: the next warp of this scheduler may start
        }
        if (active) {
            // acc <- alpha * acc (alpha = -1 was folded into the fragments), then the SAT
            // terms are added in place while the stage is still held (they read the own-row values from the ring)
            if (!alpha_m1) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) acc[mt][0] *= p.alpha, acc[mt][1] *= p.alpha;
            }
            if (SAT == 1 && mask) {  // du_j += b_j (u_j - g_j); the chunk's own rows are inside its window
                const int q0 = r0 / BX, q1 = (r0 + 1) / BX;
                const int o0 = ring_off(((q0 - blo) << 8) | (r0 - q0 * BX));
                const int o1 = ring_off(((q1 - blo) << 8) | (r0 + 1 - q1 * BX));
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    if (p.g_stride != 0) {  // per-instance boundary data
                        const int il = i0 + 8 * mt + g;
                        const long long row = b0 + (il < nb ? il : 0);
                        if (mask & 1) gs0 = __ldg(p.g + (paired ? 2 * row + h0 : row) * p.g_stride + n0);
                        if (mask & 2) gs1 = __ldg(p.g + (paired ? 2 * row + h1 : row) * p.g_stride + n1);
                    }
                    if (mask & 1) acc[mt][0] += sb.x * (su[mt * MTS + o0] - gs0);
                    if (mask & 2) acc[mt][1] += sb.y * (su[mt * MTS + o1] - gs1);
                }
            }
            if (SAT == 2 && m.x.z != 0) {
                // 1D SATs at nodes 0 and Nh-1 of every instance (the record flags the block rows that carry one).  One
                // division per instance: lane l evaluates the term for the warp's tensor rows l (and l + 32) in parallel,
                // the lane that owns the row fetches it by shuffle.
                const int flags = m.x.z;
#pragma unroll 1
                for (int k = 0; k < 8; k++) {
                    const bool left = (flags >> k) & 1, right = (flags >> (8 + k)) & 1;
                    if (!(left | right)) continue;
                    const int r = (k < 4 ? rbm.z : rbm.w - 4) + k;
                    const int h = (paired && r >= Nh) ? 1 : 0, q = r / BX;
                    const int o = ring_off(((q - blo) << 8) | (r - q * BX));
                    double sv2[(8 * MTW + 31) / 32];
#pragma unroll
                    for (int kk = 0; kk < (8 * MTW + 31) / 32; kk++) {
                        const int il = lane + 32 * kk;  // tensor row within the warp's share
                        sv2[kk] = 0.0;
                        if (il < 8 * MTW && i0 + il < nb) {
                            const long long row = (long long)b0 + i0 + il, inst = paired ? 2 * row + h : row;
                            const double u = s_u[(i0 + il) * BX + o];
                            double lb = sat_lb, rbv = sat_rb;
                            if (p.sat.bc_stride != 0) {  // per-instance boundary data
                                if (left) lb = __ldg(p.sat.bc + inst * p.sat.bc_stride);
                                if (right) rbv = __ldg(p.sat.bc + inst * p.sat.bc_stride + 1);
                            }
                            // (f - a u) / w as in sat1d_left/right, with the division replaced by the reciprocal weight
                            double v = 0.0;
                            if (left) {
                                const double t0 = p.sat.a0 * u;
                                v += ((p.sat.a0 > 0.0 ? p.sat.a0 * lb : t0) - t0) * sat_iw0;
                            }
                            if (right) {
                                const double tN = p.sat.aN * u;
                                v -= ((p.sat.aN > 0.0 ? tN : p.sat.aN * rbv) - tN) * sat_iwN;
                            }
                            sv2[kk] = v;
                        }
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        const int il = 8 * mt + g;
                        const double v = __shfl_sync(0xffffffffu, sv2[il >> 5], il & 31);
                        if (t == (k >> 1)) {
                            if (k & 1) acc[mt][1] += v;
                            else acc[mt][0] += v;
                        }
                    }
                }
            }
        }
        if (FUSED && active && r0 < N) {
            // Runge-Kutta stage update on the finished RHS: y <- ca * u + cb * extra + ch * y (u: own rows from the ring)
            const int q0 = r0 / BX, q1 = (r0 + 1) / BX;
            const int o0 = ring_off(((q0 - blo) << 8) | (r0 - q0 * BX));
            const int o1 = ring_off(((q1 - blo) << 8) | (r0 + 1 - q1 * BX));
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                double2 e = make_double2(0.0, 0.0);
                if (with_extra) {
                    if (PREFETCH_EXTRA) {
                        e = ex[PREFETCH_EXTRA ? mt : 0];
                    } else {
                        const int il = i0 + 8 * mt + g;
                        e = *reinterpret_cast<const double2*>(p.extra + (long long)(b0 + (il < nb ? il : 0)) * N + r0);
                    }
                }
                const double e0 = p.cb * e.x, e1 = p.cb * e.y;
                acc[mt][0] = fma(p.ch, acc[mt][0], fma(p.ca, su[mt * MTS + o0], e0));
                acc[mt][1] = fma(p.ch, acc[mt][1], fma(p.ca, su[mt * MTS + o1], e1));
            }
        }
        TR(6)  // scale + SAT (+ stage update)
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[st]);  // the stage (and the ring slots behind it) may be refilled
        TR(7)  // release
        if (OUT == 2) {
            // ---- epilogue through TMA: lane (g,t) deposits rows r0, r0+1 of instance 8mt + g in the warp's staging
            //      buffer; one elected lane stores quad A and quad B as {4 rows x 8*MTW instances} boxes ----
            double* so = s_out + (size_t)warp * (2 * 8 * MTW * 4);
            if (lane == 0) bulk_wait_read<0>();  // the previous boxes of this warp have been read out of the buffer
            __syncwarp();
            TR(8)  // staging buffer free
            if (active) {
                double* sq = so + (t >> 1) * (8 * MTW * 4) + g * 4 + (t & 1) * 2;
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
                    *reinterpret_cast<double2*>(sq + mt * 32) = make_double2(acc[mt][0], acc[mt][1]);
                fence_proxy_async();  // generic-proxy writes -> visible to the TMA (async proxy)
                __syncwarp();
                TR(9)  // results staged
                if (lane == 0) {
                    const int y = b0 + i0;
                    if (rbm.z < N) tma_store_2d(&tmap_out, rbm.z, y, so);
                    if (rbm.w < N) tma_store_2d(&tmap_out, rbm.w, y, so + 8 * MTW * 4);
                    bulk_commit();
                }
            }
        } else if (active && r0 < N) {
            // ---- epilogue: lane owns rows r0, r0+1 of instance i0 + 8mt + g ----
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const int il = i0 + 8 * mt + g;
                if (il >= nb) continue;
                double y0 = acc[mt][0], y1 = acc[mt][1];
                double* yp = p.dU + (long long)(b0 + il) * N + r0;
                if (VEC) {  // N even: row r0 + 1 exists
                    if (p.beta != 0.0) {
                        const double2 o = *reinterpret_cast<const double2*>(yp);
                        y0 = fma(p.beta, o.x, y0);
                        y1 = fma(p.beta, o.y, y1);
                    }
                    stg_stream2(yp, y0, y1);
                } else {
                    if (p.beta != 0.0) y0 = fma(p.beta, yp[0], y0);
                    stg_stream1(yp, y0);
                    if (r0 + 1 < N) {
                        if (p.beta != 0.0) y1 = fma(p.beta, yp[1], y1);
                        stg_stream1(yp + 1, y1);
                    }
                }
            }
        }
        m = mn;
        TR(10)  // tensor stores
#ifdef SBP_BSR_TRACE
        tr_items++;
#endif
    }
#ifdef SBP_BSR_TRACE
    if (blockIdx.x == 0 && lane == 0)
        printf("bsr trace warp %d items %d cycles/item: meta %lld consts %lld wait_full %lld first_ops %lld wait_token %lld dmma %lld sat %lld "
               "release %lld wait_staging %lld stage %lld tma_store %lld\n",
               warp, tr_items, tr_acc[0] / tr_items, tr_acc[1] / tr_items, tr_acc[2] / tr_items, tr_acc[3] / tr_items,
               tr_acc[4] / tr_items, tr_acc[5] / tr_items, tr_acc[6] / tr_items, tr_acc[7] / tr_items, tr_acc[8] / tr_items,
               tr_acc[9] / tr_items, tr_acc[10] / tr_items);
#endif
    if (PING && (warp >> 2) == 0) named_bar_sync(pp_me, 64);  // absorb the last token
    if (OUT == 2 && lane == 0) bulk_wait<0>();  // all output boxes of this warp have been written
}

static size_t bsr_smem(const sbp_op* op, int MT, int D) {
    const size_t nch = (op->bsr_nrb + kBsrCH - 1) / kBsrCH;
    (void)nch;
    const size_t ops = ((size_t)(D + 1) * op->bsr_maxf * (32 * 8 + 4 * 4) + 127) & ~(size_t)127;
    const size_t staging = (size_t)kBsrGroups * kBsrCH * 2 * 8 * MT * 4 * 8;  // per consumer warp: 2 quads x instances x 4 rows
    return (size_t)op->bsr_nbox[D] * 8 * MT * kBsrBX * 8 + ops + staging;
}
static bool bsr_fits(const sbp_ctx* ctx, const sbp_op* op, int MT, int D) {
    return bsr_smem(op, MT, D) + 1024 <= (size_t)ctx->smem_optin;
}

// true when the block-sparse tensor-core kernel can run this operator: one m-tile fits in shared memory and the batch
// rows can be described by a TMA tensor map (row pitch N*8 bytes must be a multiple of 16)
bool bsr_available(const sbp_ctx* ctx, const sbp_op* op) { return op->d_bsr_val != nullptr && bsr_fits(ctx, op, 1, 0); }
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

// 2-D tensor map over a row-major [B][N] array of doubles with boxes of box_nodes x box_inst (also used by dense_gemm.cu)
// A tensor map is a pure function of (base address, N, B, box): time steppers launch on the same few batches thousands of
// times, so the last encodings are kept (per host thread, 16 entries, round robin) instead of re-encoding 2-3 maps per launch.
int32_t make_tmap(CUtensorMap* tm, const double* base, int N, long long B, int box_nodes, int box_inst) {
    struct Entry {
        const double* base;
        long long B;
        int N, bn, bi;
        CUtensorMap map;
    };
    static thread_local Entry cache[16];
    static thread_local int used = 0, next = 0;
    for (int k = 0; k < used; k++) {
        const Entry& e = cache[k];
        if (e.base == base && e.B == B && e.N == N && e.bn == box_nodes && e.bi == box_inst) {
            *tm = e.map;
            return SBP_OK;
        }
    }
    PFN_tmapEncodeTiled enc = tmap_encoder();
    SBP_REQUIRE(enc != nullptr, SBP_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)B};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 8};
    const cuuint32_t box[2] = {(cuuint32_t)box_nodes, (cuuint32_t)box_inst};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SBP_REQUIRE(r == CUDA_SUCCESS, SBP_ECUDA, "cuTensorMapEncodeTiled failed (%d) for N=%d B=%lld", (int)r, N, B);
    Entry& e = cache[next];
    e.base = base, e.B = B, e.N = N, e.bn = box_nodes, e.bi = box_inst, e.map = *tm;
    next = (next + 1) % 16;
    used = std::min(used + 1, 16);
    return SBP_OK;
}

template <int MTW, int MH, int SAT, int OUT, bool FUSED = false>
static int32_t launch_bsr_k(sbp_ctx* ctx, const BsrParams& p, size_t smem) {
    auto kern = k_bsr<MTW, MH, SAT, OUT, FUSED>;
    constexpr int TI = 8 * MTW * MH;
    constexpr int THREADS = (kBsrGroups * kBsrCH * MH + 2) * 32;
    // the batch as a 2-D tensor [B][N] of doubles; input boxes = kBsrBX nodes x TI instances, output boxes = one quad
    // of rows x the 8*MTW instances of a consumer warp
    CUtensorMap tmap, tmap_out;
    int32_t rc = make_tmap(&tmap, p.U, p.N, p.B, kBsrBX, TI);
    if (rc) return rc;
    rc = OUT == 2 ? make_tmap(&tmap_out, p.dU, p.N, p.B, 4, 8 * MTW) : (tmap_out = tmap, SBP_OK);
    if (rc) return rc;
    SBP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SBP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const long long ntiles = (p.B + TI - 1) / TI;
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)ctx->sm_count * per_sm));
    if (ctx->tune[T_BSR_VERBOSE]) fprintf(stderr, "k_bsr<%d,%d,%d,%d>: %d CTAs/SM, grid %d, %lld tiles\n", MTW, MH, SAT, OUT, per_sm, grid, ntiles);
    kern<<<grid, THREADS, smem, ctx->stream>>>(p, tmap, tmap_out);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

template <int MTW, int MH>
static int32_t launch_bsr_mt(sbp_ctx* ctx, const BsrParams& p, size_t smem, int sat, int out) {
#define SBP_BSR_OUT(S)                                                       \
    switch (out) {                                                           \
        case 2: return launch_bsr_k<MTW, MH, S, 2>(ctx, p, smem);            \
        case 1: return launch_bsr_k<MTW, MH, S, 1>(ctx, p, smem);            \
        default: return launch_bsr_k<MTW, MH, S, 0>(ctx, p, smem);           \
    }
    if (p.fused) {  // stage updates: RHS kernels (2D or 1D SATs) with TMA output, one warp per row block
        if constexpr (MH == 1) {
            if (out == 2 && sat == 1) return launch_bsr_k<MTW, MH, 1, 2, true>(ctx, p, smem);
            if (out == 2 && sat == 2) return launch_bsr_k<MTW, MH, 2, 2, true>(ctx, p, smem);
        }
        SBP_REQUIRE(false, SBP_EUNSUPPORTED, "fused stage update needs an RHS launch with TMA output");
    }
    if (sat == 1) { SBP_BSR_OUT(1) }
    if (sat == 2) { SBP_BSR_OUT(2) }
    SBP_BSR_OUT(0)
#undef SBP_BSR_OUT
}

int32_t launch_sparse_ex(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                         double beta, const sbp_op* adv, const double* g, int64_t g_stride, const double* colscale,
                         const Sat1D* sat);

int32_t launch_bsr(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha, double beta,
                   const sbp_op* adv, const double* g, int64_t g_stride, const Sat1D* sat, const StageUpdate* su) {
    SBP_REQUIRE(B < (1ll << 31) - (1 << 20), SBP_EUNSUPPORTED, "block-sparse kernel: batch size must be < 2^31 - 2^20");
    BsrParams p;
    p.U = U;
    p.dU = dU;
    p.val = op->d_bsr_val;
    p.pos = op->d_bsr_pos;
    p.rbmeta = reinterpret_cast<const int4*>(op->d_bsr_rbmeta);
    p.chmeta = reinterpret_cast<const int4*>(op->d_bsr_chmeta);
    p.N = op->bsr_N;  // row length of the tensor view: N, or 2N for odd N (instance pairs)
    p.Nh = op->N;
    p.nrb = op->bsr_nrb;
    p.nch = (p.nrb + kBsrCH - 1) / kBsrCH;
    p.nbt = (op->bsr_N + kBsrBX - 1) / kBsrBX;
    p.maxf = op->bsr_maxf;
    p.B = B;
    p.alpha = alpha * op->scale;
    p.beta = beta;
    p.rb_bdiag = adv ? adv->d_rb_bdiag : nullptr;
    if (adv && adv->d_bsr_val_sat && adv->inner == op) {  // (launch_bsr3 decides whether the folded image is used)
        p.val_sat = adv->d_bsr_val_sat, p.val_stage = adv->d_bsr_val_stage, p.diagmask = adv->d_bsr_diagmask;
        p.nvals = (long long)op->bsr_frags * 32;
    }
    p.g = g;
    p.g_stride = g_stride;
    if (sat) p.sat = *sat;
    if (su) {
        SBP_REQUIRE((reinterpret_cast<uintptr_t>(su->extra) & 15) == 0, SBP_EUNSUPPORTED, "stage update: extra must be 16-byte aligned");
        p.fused = 1, p.extra = su->extra, p.ca = su->ca, p.cb = su->extra ? su->cb : 0.0, p.ch = su->ch;
    }
    const int satmode = (adv && adv->d_rb_bdiag) ? 1 : (sat && sat->enabled ? 2 : 0);
    // output path: TMA tensor stores when dU is 16-byte aligned and beta == 0 (SBP_BSR_OUT=1 forces plain stores)
    const int forced_out = ctx->tune[T_BSR_OUT];
    const bool aligned = (reinterpret_cast<uintptr_t>(dU) & 15) == 0;
    const bool odd = op->bsr_N != op->N;
    const int vec = !aligned ? 0 : (beta == 0.0 && forced_out == 2) ? 2 : 1;
    if (odd) {
        // pair-row view [B/2][2N]: an odd last instance has no partner row and takes the scalar kernel
        if (B % 2) {
            const long long lastoff = (B - 1) * (long long)op->N;
            int32_t rc1;
            double* out1 = dU + lastoff;
            if (su) {  // stage update: RHS of the last instance into scratch (dU may alias su->extra), combined below
                rc1 = ensure_scratch(ctx, (size_t)op->N);
                if (rc1) return rc1;
                out1 = ctx->scratch;
            }
            if (sat && sat->enabled) {
                Sat1D s1 = *sat;
                s1.bc = sat->bc + (B - 1) * sat->bc_stride;
                rc1 = launch_sparse_ex(ctx, op, out1, U + lastoff, 1, alpha, beta, nullptr, nullptr, 0, nullptr, &s1);
            } else {
                rc1 = launch_sparse(ctx, op, out1, U + lastoff, 1, alpha, beta, adv,
                                    g ? g + (B - 1) * g_stride : nullptr, g_stride, nullptr);
            }
            if (!rc1 && su)
                rc1 = launch_lincomb(ctx, dU + lastoff, U + lastoff, su->extra ? su->extra + lastoff : nullptr, out1, op->N,
                                     su->ca, 0.0, su->extra ? su->cb : 0.0, su->ch);
            if (rc1) return rc1;
            B -= 1;
            if (B == 0) return SBP_OK;
        }
        B /= 2;  // tensor rows = instance pairs
    }
    p.B = B;
    {
        int32_t rc3 = SBP_OK;
        if (launch_bsr3(ctx, op, p, satmode, vec, &rc3)) return rc3;  // lean consumer loop (bsr3.cu) when it applies
    }
    // (instances per tile, pipeline depth): the largest tile that still allows one chunk of lookahead (operator
    // fragments are re-read from L2 once per tile), then the deepest pipeline that fits; small batches use smaller
    // tiles so that every SM gets work.  SBP_BSR_MT / SBP_BSR_DEPTH override (tuning).
    const int forced_mt = ctx->tune[T_BSR_MT], forced_d = ctx->tune[T_BSR_DEPTH];
    int MT = 8;
    auto smaller = [](int mt) { return mt > 4 ? mt - 1 : mt >> 1; };
    while (MT > 1 && !bsr_fits(ctx, op, MT, 1)) MT = smaller(MT);
    while (MT > 1 && (B + 8 * MT - 1) / (8 * MT) < ctx->sm_count) MT = smaller(MT);
    if ((forced_mt == 1 || forced_mt == 2 || forced_mt == 4 || forced_mt == 3 || forced_mt == 5 || forced_mt == 6 || forced_mt == 7 || forced_mt == 8) &&
        bsr_fits(ctx, op, forced_mt, 0))
        MT = forced_mt;
    int D = 0;
    for (int d = 1; d < kBsrMaxStages; d++)
        if (d <= forced_d && bsr_fits(ctx, op, MT, d)) D = d;
    SBP_REQUIRE(bsr_fits(ctx, op, MT, D), SBP_EUNSUPPORTED, "block-sparse kernel needs %zu bytes of shared memory",
                bsr_smem(op, MT, D));
    p.nbox = op->bsr_nbox[D];
    p.nst = D + 1;
    p.item = reinterpret_cast<const int4*>(op->d_bsr_item) + (size_t)D * p.nch * kBsrCH * 2;
    const size_t smem = bsr_smem(op, MT, D);
    const bool verbose = ctx->tune[T_BSR_VERBOSE] != 0;
    if (verbose)
        fprintf(stderr, "k_bsr: N=%d B=%lld nrb=%d nch=%d frags=%lld maxf=%d MT=%d stages=%d nbox=%d smem=%zu sat=%d out=%d\n", p.N,
                (long long)B, p.nrb, p.nch, (long long)op->bsr_frags, p.maxf, MT, D + 1, p.nbox, smem, satmode, vec);
    // one consumer warp per row block of the chunk, MT m-tiles each (SBP_BSR_MH=2: two warps per row block, MT/2 each)
    const int forced_mh = ctx->tune[T_BSR_MH];
    if (forced_mh == 2 && MT >= 2) {
        switch (MT) {
            case 8: return launch_bsr_mt<4, 2>(ctx, p, smem, satmode, vec);
            case 6: return launch_bsr_mt<3, 2>(ctx, p, smem, satmode, vec);
            case 4: return launch_bsr_mt<2, 2>(ctx, p, smem, satmode, vec);
            default: return launch_bsr_mt<1, 2>(ctx, p, smem, satmode, vec);
        }
    }
    switch (MT) {
        case 8: return launch_bsr_mt<8, 1>(ctx, p, smem, satmode, vec);
        case 7: return launch_bsr_mt<7, 1>(ctx, p, smem, satmode, vec);
        case 6: return launch_bsr_mt<6, 1>(ctx, p, smem, satmode, vec);
        case 5: return launch_bsr_mt<5, 1>(ctx, p, smem, satmode, vec);
        case 3: return launch_bsr_mt<3, 1>(ctx, p, smem, satmode, vec);
        case 4: return launch_bsr_mt<4, 1>(ctx, p, smem, satmode, vec);
        case 2: return launch_bsr_mt<2, 1>(ctx, p, smem, satmode, vec);
        default: return launch_bsr_mt<1, 1>(ctx, p, smem, satmode, vec);
    }
}

}  // namespace sbp
