// bsr3.cu -- K2'/K3', second generation of the block-sparse DMMA kernel (bsr.cu): default for RHS launches and odd-N pair
// views (launch_bsr3 below; SBP_BSR_V3=0 switches it off), k_bsr keeps the plain even-N applies and the small batches.
// Same operator image, data movement (TMA box ring, operator fragments per chunk by bulk copy, TMA tensor stores) and
// barrier protocol as k_bsr, with a leaner consumer loop and the tooling used to find out what bounds the kernel.
//
// Reference math as in bsr.cu: y = alpha*scale*D u + beta*y and du = -A u + B (u - tmp1)
// (src/conservation_laws/multidimensional_linear_advection.jl:60-81), 1D SATs of the upstream semidiscretization.
//
// Differences to k_bsr:
//  * ring addresses of the A-operand gathers come precomputed from the host as BYTE offsets for the launch's tile size
//    (sbp_op::d_bsr_pos3): per fragment one add, one unsigned min (wrap) and one add instead of shift/mask/compare/
//    select/multiply-add; the chain is unrolled by two fragments over two explicit A-operand register sets, operands are
//    fetched by ld.shared at 32-bit addresses with compile-time m-tile offsets, one fragment ahead;
//  * loop invariants that the compiler rematerialised at every use are pinned in registers (also barrier addresses, lane,
//    record pointer, comparisons of double parameters), the warp index is made provably uniform (TMA stores without the
//    per-lane vote loops), item records are fetched one item ahead;
//  * box stages and operator-fragment stages are decoupled (p.nst box stages, two fragment stages);
//  * four store-helper warps (HELP) take the staging buffers from the consumer warps through mbarriers and issue the TMA
//    stores: the consumer warps neither issue stores nor wait for their buffers (config 3: 0.3256 -> 0.311 ms);
//  * round 2: L2 prefetch of the box stream (cp.async.bulk.prefetch.tensor.2d.L2), proxy fence on the helper side, a box
//    producer that keeps the ring full across tile changes (GREEDY), blocks with adjacent quads staged as [instance][8 rows]
//    and stored by one {8 x instances} box (MERGE8): config 3 0.291 ms, config 1 batched 0.155 ms;
//  * -DSBP_B3_EXP=n what-if builds, -DSBP_B3_TRACE=1 clock64 timeline, SBP_BSR_SKEW.
// What bounds it (profiles/r2_bsr_notes.txt): the consumer loop -- without any memory traffic it needs
// 0.256 ms on config 3 (1944 cycles of DMMA issue per item and scheduler, 2700-2950 cycles measured: two free-running warps
// per scheduler, ~1000 cycles of serial non-chain work per block); all memory traffic together adds 0.035 ms.  The round-1
// reading "without output 0.166 ms = the tensor pipe" came from a what-if build in which ptxas had deleted most DMMAs (their
// accumulators were never read) and is withdrawn.  An intermediate version that fetched the operator fragments from the
// L2-resident image straight into registers (no fragment stages at all) was 30 % SLOWER: the L2 round trip under load is
// exposed once per item.
#include <algorithm>
#include "bsr.cuh"

// -DSBP_B3_EXP=n: what-if builds for bottleneck hunting (WRONG RESULTS; 4 deadlocks with the store-helper warps): 1 = no DMMA, 2 = no A-operand loads inside the chain,
// 3 = no TMA box loads, 4 = no output, 5 = all tiles store to the rows of tile 0, 6 = all tiles load the rows of tile 0, 10 = no proxy fence before the TMA stores
// (7 = staging + fence but no TMA store, 8 = TMA stores without staging/fence, 9 = accumulators consumed + fence, no stores:
// measured in session 4, code removed)
#ifndef SBP_B3_EXP
#define SBP_B3_EXP 0
#endif
#ifndef SBP_B3_PING
#define SBP_B3_PING 0  // 1: the two consumer warps of a scheduler take turns on the tensor pipe (token through named barriers)
#endif
#ifndef SBP_B3_HELPER_FENCE
#define SBP_B3_HELPER_FENCE 1  // 1: the proxy fence before a TMA store is executed by the helper thread that issues the store
#endif                         // (after its acquire of the consumer warp's arrive; config 3: 0.2978 -> 0.2958 ms), 0: by the consumer warp
#ifndef SBP_B3_DEFER
#define SBP_B3_DEFER 0  // 1: deferred epilogue (see k_bsr3); measured: config 3 0.344 ms vs 0.329 without
#endif

#ifndef SBP_B3_RECPOS
#define SBP_B3_RECPOS 0  // where a consumer warp requests the record of its next row block: 0 = at the top of the item, 1 = after its chain
                         // (measured: 1 is 1-3 % slower on config 3)
#endif
#ifndef SBP_B3_GATE
#define SBP_B3_GATE 0  // k > 0: staggered chains -- a consumer warp starts its chain only while the other warp of its scheduler is not
#endif                 // inside the first k/8 of its own (shared-memory gate per scheduler); 0: free running.  Measured: k = 3..6 cost
                       // 8 % on config 3 (0.324-0.328 vs 0.299 ms in the same build), 6-13 % without any memory traffic
#ifndef SBP_B3_PRO
#define SBP_B3_PRO 0  // 1: a consumer warp steps to the next item and derives that item's addresses, masks and SAT data between
#endif                // the DMMAs of its current chain's last trip (the integer work hides behind the tensor pipe), 0: at the item top
                      // (measured: config 3 0.3175 vs 0.2913 ms, apply 0.2825 vs 0.2775; the same work placed after the chain: +3 %)
#ifndef SBP_B3_MERGE8
#define SBP_B3_MERGE8 1  // 1: a block whose quads are adjacent in memory (rows r .. r+7: banded operators) leaves as ONE {8 x instances}
#endif                   // TMA box (64-byte rows, staging stores free of bank conflicts) instead of two {4 x instances} boxes
#ifndef SBP_B3_GREEDY
#define SBP_B3_GREEDY 1  // 1: the box producer keeps the ring full -- with every item it requests all boxes of its stream (this tile and
#endif                   // the following ones) that fit behind the windows of the unreleased items, instead of just the boxes the item
                         // needs: the first window of the next tile arrives while the last chunks of this one are multiplied
#ifndef SBP_B3_SATFOLD
#define SBP_B3_SATFOLD 1  // 1: 2D RHS / stage launches of the helper kernel run on the image of A - diag(b) (stage updates: ch (A - B) - ca I,
#endif                    // rebuilt by k_stage_image per launch): no ring loads and one FP64 instruction less per SAT entry, no
                          // ca * u term and no scaling by ch after the chain
#ifndef SBP_B3_XPF
#define SBP_B3_XPF 0  // 1: the store helpers prefetch the extra operand of a stage update into L2 two items ahead
                      // (measured: stage with extra operand 0.355 -> 0.442 ms on config 3 -- the prefetches of quads compete with the box stream)
#endif
#ifndef SBP_B3_PAIR8
#define SBP_B3_PAIR8 1  // 1: x-neighbouring blocks 2j, 2j + 1 of a chunk (record flag) are staged side by side, [instance][8 nodes] per
#endif                  // quad row, and stored as two {8 x instances} boxes instead of four {4 x instances}: the what-if "same bytes in
                        // half the store boxes" was worth 12 % on config 3 (0.2826 -> 0.2487 ms)
#ifndef SBP_B3_PARKED
#define SBP_B3_PARKED 0  // 1: producers and store helpers wait with a suspend-time hint instead of tight polling (measured: no effect)
#endif
#if SBP_B3_PARKED
#define B3_PARK(bar, par) mbar_wait_parked(bar, par)
#else
#define B3_PARK(bar, par) mbar_wait(bar, par)
#endif
#ifndef SBP_B3_TRACE
#define SBP_B3_TRACE 0  // 1: clock64 stamps of CTA 0 into a global buffer (sbp_debug_set_trace), tools/trace_bsr3.py draws the timeline
#endif

namespace sbp {

#if SBP_B3_TRACE
__device__ long long* g_b3_trace = nullptr;  // [32 warps][kTraceItems][8 stamps]
constexpr int kTraceItems = 96;
#define B3_TR(w, it, idx)                                                                               \
    do {                                                                                                \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && g_b3_trace && (it) < kTraceItems)            \
            g_b3_trace[((w) * kTraceItems + (it)) * 8 + (idx)] = clock64();                             \
    } while (0)
#else
#define B3_TR(w, it, idx) do { } while (0)
#endif

#if SBP_B3_EXP == 1
#define B3_DMMA(c0, c1, a, b) asm volatile("" ::"d"(a), "d"(b))
#else
#define B3_DMMA(c0, c1, a, b) dmma884(c0, c1, a, b)
#endif

constexpr int kB3Stages = 4;      // box stages (mbarrier pairs)
#ifndef SBP_B3_FSTAGES
#define SBP_B3_FSTAGES 2
#endif
constexpr int kB3FragStages = SBP_B3_FSTAGES;  // operator-fragment stages
#ifndef SBP_B3_MERGED
#define SBP_B3_MERGED 0  // 1: boxes and operator fragments of an item complete ONE full barrier (both producers arrive on it) and
#endif                   // share the empty barrier: one mbarrier wait per item instead of two.  Measured: config 3 0.2898 either way.
constexpr bool kB3Merged = SBP_B3_MERGED != 0;
// what-ifs that keep the tensor work (ptxas deletes an mma.sync whose accumulators are never read, "asm volatile" or not: the
// "no output" build 4 measures a kernel with 2/7 of its DMMAs; check `cuobjdump -sass | grep -c DMMA` of every what-if build)
constexpr bool kNoLoads = SBP_B3_EXP == 3 || SBP_B3_EXP == 31 || SBP_B3_EXP >= 32;   // no TMA box loads
constexpr bool kNoStores = SBP_B3_EXP == 30 || SBP_B3_EXP == 31 || SBP_B3_EXP >= 32;  // helpers skip the TMA stores
constexpr bool kFragsOnce = SBP_B3_EXP == 13 || SBP_B3_EXP >= 32;                     // fragments fetched for the first tile only
constexpr bool kNoCodes = SBP_B3_EXP == 40 || SBP_B3_EXP == 41 || SBP_B3_EXP == 43;   // ring addresses do not depend on loaded codes
constexpr bool kNoOperands = SBP_B3_EXP == 2 || SBP_B3_EXP == 41 || SBP_B3_EXP == 42;  // one A operand per fragment comes from a register

// shared-memory loads at a 32-bit shared address + compile-time offset (volatile: they keep their place in the software pipeline)
template <uint32_t OFF>
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];\n" : "=d"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <uint32_t OFF>
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];\n" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int MT, uint32_t MTB, int I = 0>
__device__ __forceinline__ void lds_tiles(double (&a)[MT], uint32_t addr) {
    if constexpr (I < MT) {
        a[I] = lds64<I * MTB>(addr);
        lds_tiles<MT, MTB, I + 1>(a, addr);
    }
}

// CW consumer warps (8 or 12) share the row blocks of the item stream round-robin: block k = 8 * item + slot belongs to warp
// k % CW.  CW = 8: a warp keeps its slot and has a block in every item; CW = 12: three consumer warps per scheduler, a warp
// has a block in two out of three items (it still passes every item's "full" barrier, in order).
// HELP: four extra warps issue the TMA stores (helper h serves consumer warps h and h + 4): a consumer warp deposits its
// accumulators in its staging buffer and signals an mbarrier instead of issuing the stores and waiting for the buffer itself.
// FUSED: Runge-Kutta stage update on the finished RHS (sbp_rhs_advection*_stage), as in k_bsr<..., FUSED>.
template <int MT, int SAT, int OUT, int CW, bool HELP = false, bool FUSED = false, bool XTRA = false, bool GPI = false, bool PAIR = false>
__global__ void __launch_bounds__((CW + 2 + (HELP ? 4 : 0)) * 32)
    k_bsr3(const BsrParams p, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out,
           const __grid_constant__ CUtensorMap tmap_ex, const __grid_constant__ CUtensorMap tmap_out8) {
    // MERGE8: adjacent quads of a block are staged as [instance][8 rows] and stored by one box (not when the staging buffer also
    // carries the extra operand of a fused stage update, which arrives quad by quad)
    constexpr bool MERGE8 = SBP_B3_MERGE8 && HELP && SBP_B3_EXP == 0;
    // PAIR8: the same for two x-neighbouring blocks (not with an extra operand in the staging buffers)
    constexpr bool PAIR8 = SBP_B3_PAIR8 && PAIR && MERGE8 && kBsrCH == 8 && !(FUSED && XTRA);  // (template: kernels of operators without pairs carry none of it)
    constexpr bool VEC = OUT == 1 || OUT == 2;
    constexpr int BX = kBsrBX;
    constexpr int TI = 8 * MT;
    constexpr int CONS = CW;                   // consumer warps
    static_assert(CW >= kBsrCH && CW < 2 * kBsrCH, "a consumer warp has at most one row block per item");
    constexpr int BOX = TI * BX;               // doubles per box
    constexpr uint32_t BOXB = BOX * 8;         // bytes per box
    constexpr uint32_t MTB = 8 * BX * 8;       // byte distance between the m-tiles of a lane
    constexpr int PS = 4;                      // ring codes per fragment
    extern __shared__ __align__(128) double smem[];
    // boxes and operator fragments have separate stage counts: p.nst box stages (HBM latency: as deep as the ring allows),
    // kB3FragStages fragment stages (L2 latency: two are enough, each costs maxf * 272 bytes)
    // all mbarriers in one array: a consumer warp addresses them as (one pinned 32-bit shared address) + constant
    constexpr int O_FULL = 0, O_EMPTY = O_FULL + kB3Stages, O_FFULL = O_EMPTY + kB3Stages, O_FEMPTY = O_FFULL + kB3FragStages;
    constexpr int O_OFULL = O_FEMPTY + kB3FragStages, O_OEMPTY = O_OFULL + kBsrCH, O_XFULL = O_OEMPTY + kBsrCH;
    __shared__ uint64_t bars[O_XFULL + kBsrCH];
    uint64_t* const bar_full = bars + O_FULL;
    uint64_t* const bar_empty = bars + O_EMPTY;
    uint64_t* const fbar_full = bars + O_FFULL;
    uint64_t* const fbar_empty = bars + O_FEMPTY;
    uint64_t* const out_full = bars + O_OFULL;    // HELP: staging buffer of consumer warp w filled / stored
    uint64_t* const out_empty = bars + O_OEMPTY;
    uint64_t* const ext_full = bars + O_XFULL;    // HELP + FUSED: the quads of `extra` for the warp's next block are in its staging buffer
    __shared__ int gate[4];     // SBP_B3_GATE: id of the warp inside the gated part of its chain, -1: open
    __shared__ int sem_fin[4];  // CW == 12, p.sem: chains finished per scheduler (at most two of its three warps multiply at a time)
    // XT: the extra operand of a fused stage update travels through the staging buffers (TMA loads issued by the helpers)
    // (XTRA: the launch has an extra operand with cb != 0 -- a template parameter, so that stage updates without one carry none of
    // the machinery: the runtime form of this flag cost 2-7 % in the same way as a runtime SAT fold, profiles/r2_bsr_notes.txt 8)
    constexpr bool XT = HELP && FUSED && XTRA;
    static_assert(!HELP || (CW == kBsrCH && OUT == 2 && kBsrCH % 4 == 0), "helper warps: one consumer warp per row block of a chunk, TMA output");
    double* s_u = smem;                                                              // nbox * BOX
    double* s_val = s_u + (size_t)p.nbox * BOX;                                      // nst * maxf * 32
    const int fs = kB3Merged ? p.nst : kB3FragStages;                                 // fragment stages
    int* s_pos = reinterpret_cast<int*>(s_val + (size_t)fs * p.maxf * 32);            // fs * maxf * PS
    // output staging, per consumer warp: [quad A | quad B][8 * MT instances][4 rows]  (128-byte aligned)
    double* s_out = reinterpret_cast<double*>(s_pos + (((size_t)fs * p.maxf * PS + 31) & ~(size_t)31));
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
    const int N = p.N, nst = p.nst, nbox = p.nbox;
    const long long ntiles = (p.B + TI - 1) / TI;
    const int stepb = p.nbt % nbox;  // ring offset advance per tile (in boxes)
    const bool alpha_m1 = p.alpha == -1.0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nst; s++) {
            mbar_init(&bar_full[s], kB3Merged ? 2 : 1);  // the box producer's arrive.expect_tx (+ the fragment producer's)
            mbar_init(&bar_empty[s], CONS);  // lane 0 of EVERY consumer warp, block in this item or not: a stage that could be
        }                                    // refilled without a warp's arrival might lap that warp's phase tracking
        for (int s = 0; s < kB3FragStages; s++) {
            mbar_init(&fbar_full[s], 1);
            mbar_init(&fbar_empty[s], CONS);
        }
        for (int s = 0; s < kBsrCH; s++) {
            mbar_init(&out_full[s], 1);
            mbar_init(&out_empty[s], 1);
            mbar_init(&ext_full[s], 1);
        }
        for (int s = 0; s < 4; s++) sem_fin[s] = 0, gate[s] = -1;
        mbar_fence_init();
    }
    __syncthreads();

    if (HELP && warp >= CONS + 2) {
        // ------- store helpers: per item and served consumer warp, wait for its staging buffer, store the two quads, hand it back -------
#if SBP_B3_EXP == 11 || SBP_B3_EXP == 12
        // what-if (WRONG RESULTS): one helper thread stores the staging area of all eight consumer warps as two {32 rows x TI}
        // boxes (11) / four {16 x TI} boxes (12) per item instead of sixteen {4 x TI} boxes: same bytes, 1/8 (1/4) of the TMA rows
        if (lane == 0 && warp == CONS + 2) {
            int ph = 0;
            constexpr int W = SBP_B3_EXP == 11 ? 32 : 16;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int b0 = (int)(tile * TI);
                for (int cc = 0; cc < p.nch; cc++) {
                    for (int w = 0; w < kBsrCH; w++) mbar_wait(&out_full[w], ph);
                    const int r = min(64 * cc, N - 64) & ~1;
                    for (int k = 0; k < 64 / W; k++) tma_store_2d(&tmap_ex, r + W * k, b0, s_out + (size_t)k * W * TI);
                    bulk_commit();
                    bulk_wait_read<0>();
                    for (int w = 0; w < kBsrCH; w++) mbar_arrive(&out_empty[w]);
                    ph ^= 1;
                }
            }
            bulk_wait<0>();
        }
        return;
#endif
        if (lane == 0) {
            const int h = warp - (CONS + 2);
            int c = 0, ph = 0;
#if SBP_B3_TRACE
            int trh = 0;
#endif
            const int4* rec = p.item;
            constexpr int KH = kBsrCH / 4;  // consumer warps served by one helper
            // slots of this helper's blocks: h, h + 4, ... -- or, with pairs (p.pair8), the neighbouring slots KH h, KH h + 1, ...;
            // either way they belong to the two warps of one scheduler (see `slot` in the consumer part)
            constexpr bool pm = PAIR8;
            auto hslot = [&](int k) { return pm ? KH * h + k : h + 4 * k; };
            int rowA[KH], rowB[KH], pairf = 0;
            for (int k = 0; k < KH; k++) {
                const int4 a = __ldg(rec + 2 * hslot(k));
                rowA[k] = a.z, rowB[k] = a.w;
            }
            if (pm) pairf = __ldg(rec + 2 * (KH * h) + 1).w;  // 1: the helper's first two blocks form a pair in this item
            // XT: quads rA / rB of `extra` for tensor rows y.. of consumer warp w -> its staging buffer, completion on ext_full[w]
            auto load_extra = [&](int w, int rA, int rB, int y) {
                double* so = s_out + (size_t)w * (2 * 8 * MT * 4);
                const unsigned quads = (rA < N ? 1u : 0u) + (rA < N && rB < N ? 1u : 0u);
                mbar_arrive_expect_tx(&ext_full[w], quads * (unsigned)(4 * TI * 8));
                if (rA < N) tma_load_2d(so, &tmap_ex, rA, y, &ext_full[w]);
                if (rA < N && rB < N) tma_load_2d(so + 8 * MT * 4, &tmap_ex, rB, y, &ext_full[w]);
            };
            if (XT && (long long)blockIdx.x < ntiles)
                for (int k = 0; k < KH; k++) load_extra(hslot(k), rowA[k], rowB[k], (int)(blockIdx.x * TI));
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int b0 = (int)(tile * TI);
                for (int cc = 0; cc < p.nch; cc++) {
                    const int cn = c + 1 < p.nch ? c + 1 : 0;
                    int nA[KH], nB[KH], npairf = 0;
                    for (int k = 0; k < KH; k++) {  // next item's rows (the records repeat with the tile)
                        const int4 a = __ldg(rec + 2 * (kBsrCH * cn + hslot(k)));
                        nA[k] = a.z, nB[k] = a.w;
                    }
                    if (pm) npairf = __ldg(rec + 2 * (kBsrCH * cn + KH * h) + 1).w;
                    // first tensor row of the next item (same tile, or this CTA's next tile); -1: there is none
                    const long long ntile = cc + 1 < p.nch ? tile : tile + gridDim.x;
                    const int nb0 = ntile < ntiles ? (int)(ntile * TI) : -1;
                    if (XT && SBP_B3_XPF) {  // L2 prefetch of the extra operand's quads two items ahead (the loads below are one ahead)
                        const int cn2 = cn + 1 < p.nch ? cn + 1 : 0;
                        const long long t2 = cn2 == 0 ? ntile + gridDim.x : ntile;  // tile of the item after next
                        if (t2 < ntiles)
                            for (int k = 0; k < KH; k++) {
                                const int4 a = __ldg(rec + 2 * (kBsrCH * cn2 + hslot(k)));
                                if (a.z < N)
                                    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(&tmap_ex),
                                                 "r"(a.z), "r"((int)(t2 * TI))
                                                 : "memory");
                                if (a.z < N && a.w < N)
                                    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(&tmap_ex),
                                                 "r"(a.w), "r"((int)(t2 * TI))
                                                 : "memory");
                            }
                    }
                    if (pm && KH == 2 && pairf == 1) {
                        // a pair: both warps have staged [instance][8 nodes] for quad row A and for quad row B in their common region
                        const int w = KH * h;
                        B3_PARK(&out_full[w], ph);
                        B3_PARK(&out_full[w + 1], ph);
                        if (SBP_B3_HELPER_FENCE) fence_proxy_async();
                        const double* so = s_out + (size_t)w * (2 * 8 * MT * 4);
                        if (!kNoStores) {
                            tma_store_2d(&tmap_out8, rowA[0], b0, so);
                            tma_store_2d(&tmap_out8, rowB[0], b0, so + 8 * MT * 8);
                            bulk_commit();
                            bulk_wait_read<0>();
                        }
                        mbar_arrive(&out_empty[w]);
                        mbar_arrive(&out_empty[w + 1]);
                    } else
                    for (int k = 0; k < KH; k++) {
                        const int w = hslot(k);
                        B3_PARK(&out_full[w], ph);  // (acquire: the consumer warp's staging stores are visible)
                        B3_TR(warp, trh, 2 * k);
                        if (SBP_B3_HELPER_FENCE) fence_proxy_async();  // generic-proxy writes -> async proxy, on the issuing side
                        const double* so = s_out + (size_t)w * (2 * 8 * MT * 4);
                        if (rowA[k] < N && !kNoStores) {
                            if (MERGE8 && !XT && rowB[k] == rowA[k] + 4) {
                                tma_store_2d(&tmap_out8, rowA[k], b0, so);
                            } else {
                                tma_store_2d(&tmap_out, rowA[k], b0, so);
                                if (rowB[k] < N) tma_store_2d(&tmap_out, rowB[k], b0, so + 8 * MT * 4);
                            }
                            bulk_commit();
                            bulk_wait_read<0>();
                        }
                        if (XT && nb0 >= 0) load_extra(w, nA[k], nB[k], nb0);  // the buffer is free: fetch the next block's operand
                        mbar_arrive(&out_empty[w]);
                        B3_TR(warp, trh, 2 * k + 1);
                    }
#if SBP_B3_TRACE
                    trh++;
#endif
                    for (int k = 0; k < KH; k++) rowA[k] = nA[k], rowB[k] = nB[k];
                    pairf = npairf;
                    c = cn;
                    ph ^= 1;
                }
            }
            bulk_wait<0>();
        }
        return;
    }
    if (warp >= CONS) {
        // ------- producers (as in bsr.cu): warp CONS requests the boxes of the batch, warp CONS + 1 the operator fragments -------
        if (lane == 0) {
            const bool boxes = warp == CONS;
            const int pst = (boxes || kB3Merged) ? nst : kB3FragStages;  // stages of this producer's barrier set
            uint64_t* const pfull = (boxes || kB3Merged) ? bar_full : fbar_full;
            uint64_t* const pempty = (boxes || kB3Merged) ? bar_empty : fbar_empty;
            int st = 0, round = 0, slot = 0;
#if SBP_B3_TRACE
            int trp = 0;
#endif
            const uint32_t s_u32 = smem_u32(s_u), s_val32 = smem_u32(s_val), s_pos32 = smem_u32(s_pos);
            const uint32_t full32 = smem_u32(pfull);
            int4 cm = __ldg(p.chmeta);
            // GREEDY: next box of this CTA's stream (tile rtile, box rj), boxes requested counted from the current tile's box 0,
            // boxes left in the stream, and the window bases of the last items (relative to the current tile)
            long long rtile = blockIdx.x, req = 0;
            long long stream_left = (long long)blockIdx.x < ntiles ? ((ntiles - 1 - blockIdx.x) / gridDim.x + 1) * p.nbt : 0;
            int rj = 0;
            int lo_hist[kB3Stages] = {};
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int b0 = (int)(tile * TI);
                int jb = 0, x = 0;  // boxes of this tile already requested, node coordinate of the next box
                for (int c = 0; c < p.nch; c++) {
                    const int4 cmn = __ldg(p.chmeta + (c + 1 < p.nch ? c + 1 : 0));
                    if (round > 0) B3_PARK(&pempty[st], (round - 1) & 1);
                    B3_TR(warp, trp, 0);
                    const uint32_t bar = full32 + 8u * st;
                    if (boxes && SBP_B3_GREEDY && !kNoLoads) {
                        // lowest box (relative to this tile) that an unreleased item may still read: item s - D, D = nst - 1
                        for (int k = kB3Stages - 1; k > 0; k--) lo_hist[k] = lo_hist[k - 1];
                        lo_hist[0] = cm.z;
                        const long long fit = (long long)lo_hist[nst - 1] + nbox - req;  // free ring slots
                        const int n = (int)max(0LL, min(fit, stream_left));
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"((unsigned)n * BOXB)
                                     : "memory");
                        for (int k = 0; k < n; k++) {
                            if (p.pf > 0) {
                                int jp = rj + p.pf;
                                long long tp = rtile;
                                if (jp >= p.nbt) jp -= p.nbt, tp += gridDim.x;
                                if (tp < ntiles)
                                    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(&tmap),
                                                 "r"(jp * BX), "r"((int)(tp * TI))
                                                 : "memory");
                            }
                            const uint32_t dst = s_u32 + (uint32_t)slot * BOXB;
                            asm volatile(
                                "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, "
                                "{%2, %3}], [%4];\n" ::"r"(dst),
                                "l"(&tmap), "r"(rj * BX), "r"((int)(rtile * TI)), "r"(bar)
                                : "memory");
                            if (++slot == nbox) slot = 0;
                            if (++rj == p.nbt) rj = 0, rtile += gridDim.x;
                        }
                        req += n, stream_left -= n;
                    } else if (boxes) {
                        const int jn = cm.w;
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar),
                                     "r"(kNoLoads ? 0u : (unsigned)(jn - jb) * BOXB)
                                     : "memory");
                        if (kNoLoads) jb = jn;
                        for (; jb < jn; jb++, x += BX) {
                            if (p.pf > 0) {  // L2 prefetch of the box p.pf positions ahead in this CTA's box stream (HBM -> L2, no
                                int jp = jb + p.pf;  // shared memory): the ring's lookahead is shallower than the HBM latency needs
                                long long tp = tile;
                                if (jp >= p.nbt) jp -= p.nbt, tp += gridDim.x;
                                if (tp < ntiles)
                                    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(&tmap),
                                                 "r"(jp * BX), "r"((int)(tp * TI))
                                                 : "memory");
                            }
                            const uint32_t dst = s_u32 + (uint32_t)slot * BOXB;
                            asm volatile(
                                "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, "
                                "{%2, %3}], [%4];\n" ::"r"(dst),
                                "l"(&tmap), "r"(x), "r"(SBP_B3_EXP == 6 ? 0 : b0), "r"(bar)
                                : "memory");
                            if (++slot == nbox) slot = 0;
                        }
                    } else {
                        // what-if 13 (WRONG RESULTS): operator fragments are fetched for the CTA's first tile only
                        const unsigned nfr = (kFragsOnce && tile != (long long)blockIdx.x) ? 0u : (unsigned)(cm.y - cm.x);
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
                    B3_TR(warp, trp, 1);
#if SBP_B3_TRACE
                    trp++;
#endif
                    if (++st == pst) st = 0, round++;
                    cm = cmn;
                }
                if (boxes && SBP_B3_GREEDY && !kNoLoads) {  // the next tile's box 0 becomes the origin
                    req -= p.nbt;
                    for (int k = 0; k < kB3Stages; k++) lo_hist[k] -= p.nbt;
                } else {
                    // boxes of the tile that no chunk needs are skipped: keep slot == (virtual box index) mod nbox
                    slot += (p.nbt - jb) % nbox;
                    if (slot >= nbox) slot -= nbox;
                }
            }
        }
        return;
    }

    // ---------------------------------------- consumers ----------------------------------------
    const int g = lane >> 2, t = lane & 3;
    const int Nh = p.Nh;
    const bool paired = Nh != N;
    const uint32_t ringB = (uint32_t)nbox * BOXB;
    const char* const su_lane = reinterpret_cast<const char*>(s_u + g * BX);  // tensor row 8 mt + g lives MTB * mt further
    // slot of this warp's row block inside a chunk.  With store helpers the two warps of a scheduler (w, w + 4) take NEIGHBOURING
    // slots (2w, 2w + 1): helper h then serves one scheduler's two warps -- which rarely stage at the same time -- and a pair of
    // x-neighbouring blocks (PAIR8) lies in their hands (serving warps of different schedulers cost 4 % on banded operators)
    const int slot = PAIR8 ? (warp & 3) * 2 + (warp >> 2) : warp;
    constexpr bool pair_on = PAIR8;
    double* const so = s_out + (size_t)slot * (2 * 8 * MT * 4);
    double* const sq = so + (t >> 1) * (8 * MT * 4) + g * 4 + (t & 1) * 2;
    double* const sq8 = so + g * 8 + 2 * t;  // MERGE8 layout
    // PAIR8 layout in the common region of warps 2j, 2j + 1: [quad row A | quad row B][instance][8 nodes], this block's nodes 4 (warp & 1)..
    double* const sqp = s_out + (size_t)(slot & ~1) * (2 * 8 * MT * 4) + (t >> 1) * (8 * MT * 8) + g * 8 + (slot & 1) * 4 + (t & 1) * 2;

    // item records (one per row block, independent of the tile): fetched two items ahead
    struct Rec {
        int4 a;      // {first fragment relative to the chunk, fragments, first row of quad A, of quad B}
        int4 b;      // {ring slot of the chunk's lowest box, lowest box, 1D SAT flags, -}
        double2 sb;  // 2D SAT coefficients of the lane's two rows (0: none)
    };
    const double2* const sbs = reinterpret_cast<const double2*>(p.rb_bdiag) + t;
    const int4* item_p = p.item;  // (pinned: one LDC.64 + its latency per item otherwise)
    asm volatile("" : "+l"(item_p));
    auto load_rec = [&](int rb) {  // rb = 8 * chunk + slot
        Rec r;
        r.a = __ldg(item_p + 2 * rb);
        r.b = __ldg(item_p + 2 * rb + 1);
        r.sb = make_double2(0.0, 0.0);
        if (SAT == 1) r.sb = __ldg(sbs + 4 * rb);
        return r;
    };

    int c = 0, st = 0, phase = 0, fst = 0, fphase = 0;
    int tile = blockIdx.x, b0 = blockIdx.x * TI;  // B < 2^31 (checked by the launcher)
    uint32_t offB = 0;                            // ring offset of the tile's box 0 (bytes)
    const uint32_t stepB = (uint32_t)stepb * BOXB;
    const int ntl = (int)ntiles, tstep = gridDim.x * TI, Bi = (int)p.B;
    // position of a warp's next row block in the (periodic) record array: chunk pc, slot ps; the one after it is CW blocks on
    const int nch_p = __shfl_sync(0xffffffffu, p.nch, 0), grid_p = __shfl_sync(0xffffffffu, (int)gridDim.x, 0);  // pinned (see below)
    int ps = slot % kBsrCH, pc = (slot / kBsrCH) % p.nch;
    auto advance = [&]() {
        const int q = ps + CW;
        ps = q % kBsrCH;
        pc += q / kBsrCH;
        while (pc >= nch_p) pc -= nch_p;
    };
    Rec r0 = load_rec(kBsrCH * pc + ps);
    advance();
    int skip = warp / kBsrCH, myslot = warp % kBsrCH;  // items to pass before this warp's next block; its slot there
    // the launcher gives SAT == 1 helper kernels the folded image (or does not take this kernel at all)
    constexpr bool satfold = SBP_B3_SATFOLD && SAT == 1 && HELP && SBP_B3_EXP == 0;
    // boundary data per instance (g_stride != 0): a template parameter (GPI) of the kernels with the folded image, a runtime test elsewhere
    const bool g_per_inst = satfold ? GPI : p.g_stride != 0;
    // stage updates of the helper kernels run on an image that contains ca I and the factor ch (2D: ch (A - B) - ca I with
    // alpha = -1; 1D: ch alpha D + ca I with alpha = +1): after the chain only cb * extra and the SAT constants are left
    constexpr bool stagefold = FUSED && HELP && SBP_B3_SATFOLD && SBP_B3_EXP == 0 && (SAT == 1 || SAT == 2);
    double sat_iw0 = 0.0, sat_iwN = 0.0, sat_lb = 0.0, sat_rb = 0.0, sat_c0 = 0.0, sat_cN = 0.0;
    // boundary data per instance: a template parameter (GPI) of the helper kernels, a runtime test elsewhere
    const bool bc_per_inst = (SAT == 2 && HELP) ? GPI : p.sat.bc_stride != 0;
    if (SAT == 2) {
        sat_iw0 = 1.0 / p.sat.w0, sat_iwN = 1.0 / p.sat.wN;
        if (p.sat.bc_stride == 0) sat_lb = __ldg(p.sat.bc), sat_rb = __ldg(p.sat.bc + 1);
        // left boundary: inflow (a0 > 0) contributes a0 (lb - u) / w0, outflow nothing; right boundary: aN < 0 contributes
        // -aN (rb - u) / wN -- the comparisons are made here, once (sat_c0 / sat_cN = coefficient of (bc - u), 0: no term)
        sat_c0 = p.sat.a0 > 0.0 ? p.sat.a0 * sat_iw0 : 0.0;
        sat_cN = p.sat.aN > 0.0 ? 0.0 : -p.sat.aN * sat_iwN;
        if (stagefold) sat_c0 *= p.ch, sat_cN *= p.ch;
    }
    // loop invariants the compiler would otherwise rematerialise at every use: pinned in registers
    int sgn = alpha_m1 ? (int)0x80000000u : 0;
    uint32_t su32 = smem_u32(su_lane), ringBr = ringB;
    uint32_t sval32 = smem_u32(s_val) + lane * 8, spos32 = smem_u32(s_pos) + t * 4;  // stage 0, this lane
    uint32_t stvalB = (uint32_t)p.maxf * 256u, stposB = (uint32_t)p.maxf * (4u * PS);
    asm volatile("" : "+r"(sgn), "+r"(su32), "+r"(ringBr), "+r"(sval32), "+r"(spos32), "+r"(stvalB), "+r"(stposB));
    // barrier addresses: the array base and this warp's entry of the per-warp barriers
    uint32_t bar32 = smem_u32(bars), wbar32 = smem_u32(bars + slot);
    int lane_p = lane;  // (pinned as well: threadIdx.x would be re-read through S2R at every `lane == 0`)
    asm volatile("" : "+r"(bar32), "+r"(wbar32), "+r"(lane_p));
    const bool lane0 = lane_p == 0;
    // (routed through a shuffle: ptxas does not rematerialise a shuffle result from the constant bank)
    auto pin = [](int v) { return __shfl_sync(0xffffffffu, v, 0); };
    const int nch = nch_p, Nr = pin(N), nstr = pin(nst), tstepr = pin(tstep), ntlr = pin(ntl), Bir = pin(Bi);
    // the image arrives scaled: alpha = +-1 (sign flip of the B operand), any other alpha through k_stage_image in front of the launch
    constexpr int scale_needed = 0;
    // comparisons of double parameters, made once: a DSETP inside the item loop queues behind the DMMAs of the other warp (ncu: 4 %
    // of the fused kernel's samples on five of them per block)
    const int cbnz = pin(p.cb != 0.0 ? 1 : 0);
    constexpr bool XTc = XT;
    // (measured and dropped: -b_j g_j as the accumulators' initial value, one DFMA per SAT entry after the chain instead of DADD +
    // DFMA: config 3 0.3095 vs 0.2929 ms; two instead of three FP64 instructions per entry in stage updates without an extra
    // operand: SSPRK53 step 1.658 vs 1.622 ms)
    const uint32_t stepBr = (uint32_t)pin((int)stepB);
    sgn = pin(sgn);
    uint32_t svst = sval32, spst = spos32;  // fragment buffers of stage st

    // PING: tensor-pipe token of this warp's scheduler, barrier id (1 + 2 * scheduler + k) = "warp k may multiply"
    constexpr bool PING = SBP_B3_PING && CW == kBsrCH;
    const int pp_me = 1 + 2 * (warp & 3) + (warp >> 2), pp_next = 1 + 2 * (warp & 3) + (1 - (warp >> 2));
    if (PING && (warp >> 2) == 1) named_bar_arrive(pp_next, 64);  // the first turn belongs to warp k = 0
    if (p.skew > 0 && warp >= 4) {  // de-phase the two consumer warps of a scheduler (chains against scalar phases)
        const long long t0 = clock64();
        while (clock64() - t0 < p.skew) {
        }
    }
    // DEFER: the results of a block leave the registers one block later, after the NEXT chain has been issued (two accumulator
    // sets, the loop is unrolled by two for static register indices): the warp never waits for its own DMMAs to complete, and
    // the staging stores, the proxy fence and the TMA store issue run while the tensor pipe works on the next chain.
    constexpr bool DEFER = SBP_B3_DEFER && OUT == 2 && CW == kBsrCH && SBP_B3_EXP == 0 && !HELP;
    int out_phase = 0, ext_phase = 0;
    bool prev_pair = false;  // PAIR8: the warp's previous block left as part of a pair
#if SBP_B3_TRACE
    int trit = 0;
#endif
    int sem_n = warp >> 2;  // CW == 12: index of this warp's next block among the blocks of its scheduler (rotation of three warps)
    double accs[DEFER ? 2 : 1][MT][2];
    int pv_rowA = 0, pv_rowB = 0, pv_b0 = 0, pv_state = 0;  // deferred block: rows of its quads, first tensor row, 1 = pending | 2 = scale | 4 = set
    // staging + TMA stores of one block's accumulators
    auto emit = [&](double (&a)[MT][2], int rowA, int rowB, int yb0, bool scale) {
        if (lane == 0) bulk_wait_read<0>();  // the previous boxes of this warp have been read out of the buffer
        __syncwarp();
        if (scale) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++) a[mt][0] *= p.alpha, a[mt][1] *= p.alpha;
        }
#pragma unroll
        for (int mt = 0; mt < MT; mt++) *reinterpret_cast<double2*>(sq + mt * 32) = make_double2(a[mt][0], a[mt][1]);
        if (SBP_B3_EXP != 10) fence_proxy_async();  // generic-proxy writes -> visible to the TMA (async proxy); what-if 10: none
        __syncwarp();
        if (lane == 0) {
            const int yo = SBP_B3_EXP == 5 ? 0 : yb0;  // what-if 5: every tile writes the rows of tile 0 (L2-resident output)
            if (rowA < Nr && !kNoStores) tma_store_2d(&tmap_out, rowA, yo, so);
            if (rowB < Nr && !kNoStores) tma_store_2d(&tmap_out, rowB, yo, so + 8 * MT * 4);
            bulk_commit();
        }
    };
    // what a warp derives from the record of a row block before it can multiply (everything but the data itself)
    struct Item {
        int nf, rA, rB, lowbox, flags, pair;  // fragments, first row of quad A / of quad B, the chunk's lowest box, 1D SAT flags, pair flag
        int r0w, mask, n0, n1, h0, h1;  // the lane's rows r0w, r0w + 1: SAT mask, node and half (instance pairs)
        uint32_t baseB, sv, sp;         // ring window base (bytes), shared addresses of the block's fragments / ring codes
        double2 sb;
        double gs0, gs1;                // boundary data shared by all instances
    };
    auto prologue = [&](const Rec& r, uint32_t offB_, uint32_t svst_, uint32_t spst_) {
        Item I;
        I.nf = r.a.y, I.rA = r.a.z, I.rB = r.a.w, I.lowbox = r.b.y, I.flags = r.b.z, I.pair = r.b.w;
        // window base of this item in the ring (bytes): slot of the chunk's lowest box in this tile
        uint32_t baseB = (uint32_t)r.b.x * BOXB + offB_;
        baseB = min(baseB, baseB - ringBr);
        asm volatile("" : "+r"(baseB));
        I.baseB = baseB;
        I.sb = r.sb;
        I.r0w = (t < 2 ? r.a.z : r.a.w - 4) + 2 * t;  // lane owns rows r0w, r0w + 1 (pair t of the block)
        // b_j != 0 tested on the bit pattern (no trip through the FP64 pipe)
        I.mask = (SAT == 1) ? ((((__double2hiint(r.sb.x) << 1) | __double2loint(r.sb.x)) != 0 ? 1 : 0) |
                               (((__double2hiint(r.sb.y) << 1) | __double2loint(r.sb.y)) != 0 ? 2 : 0))
                            : 0;
        I.h0 = (paired && I.r0w >= Nh) ? 1 : 0, I.h1 = (paired && I.r0w + 1 >= Nh) ? 1 : 0;
        I.n0 = I.r0w - I.h0 * Nh, I.n1 = I.r0w + 1 - I.h1 * Nh;
        I.gs0 = 0.0, I.gs1 = 0.0;
        if (SAT == 1 && I.mask && !g_per_inst) {
            if (I.mask & 1) I.gs0 = __ldg(p.g + I.n0);
            if (I.mask & 2) I.gs1 = __ldg(p.g + I.n1);
        }
        // this row block's fragments in the stage buffers: B operand of this lane, ring code of this lane's column
        I.sv = svst_ + (uint32_t)r.a.x * 256u, I.sp = spst_ + (uint32_t)r.a.x * (4u * PS);
        return I;
    };
    constexpr bool PRO = SBP_B3_PRO && CW == kBsrCH && SBP_B3_RECPOS == 0;
    // step to the next item: chunk / tile counters, ring offset, stage and phase of the barriers, fragment buffers
    auto next_item = [&]() {
        if (++c == nch) {
            c = 0;
            tile += grid_p;
            b0 += tstepr;
            offB += stepBr;
            offB = min(offB, offB - ringBr);
        }
        svst += stvalB, spst += stposB;
        if (++st == nstr) {
            st = 0, phase ^= 1;
            if (kB3Merged) svst = sval32, spst = spos32;
        }
        if (!kB3Merged && ++fst == kB3FragStages) fst = 0, fphase ^= 1, svst = sval32, spst = spos32;
    };
    Item nxt = prologue(r0, offB, svst, spst);
    while (true) {
#pragma unroll
     for (int hf = 0; hf < (DEFER ? 2 : 1); hf++) {
      if (tile >= ntlr) goto done;
      if (CW == kBsrCH || skip == 0) {
        B3_TR(warp, trit, 0);
#if !SBP_B3_RECPOS
        const Rec r1 = load_rec(kBsrCH * pc + ps);
        advance();
#endif
        const Item I = PRO ? nxt : prologue(r0, offB, svst, spst);
        const int nf = I.nf, r0w = I.r0w, mask = I.mask, n0 = I.n0, n1 = I.n1, h0 = I.h0, h1 = I.h1;
        const bool active = I.rA < Nr;  // padding records of the last chunk carry rows >= N and no fragments
        const uint32_t baseB = I.baseB;
        const double2 sb = I.sb;
        double gs0 = I.gs0, gs1 = I.gs1;
        // this item's first tensor row, barrier stages and phases (PRO steps on before they are used for the last time)
        const int b0c = b0, stc = st, fstc = fst;
        bool stepped = false;
        double(&acc)[MT][2] = accs[DEFER ? hf : 0];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) acc[mt][0] = acc[mt][1] = 0.0;
        // stage update: the lane's entries of `extra` are requested now and consumed after the chain
        double2 ex[(FUSED && !HELP) ? MT : 1];  // (with helper warps the operand comes through the staging buffer: XT)
        const bool with_extra = FUSED && (HELP ? XTRA : cbnz != 0) && active && r0w < Nr;
        if (FUSED && !HELP && with_extra) {
            const int nbx = min(Bir - b0, TI);
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const int il = 8 * mt + g;
                ex[(FUSED && !HELP) ? mt : 0] = *reinterpret_cast<const double2*>(p.extra + (long long)(b0 + (il < nbx ? il : 0)) * N + r0w);
            }
        }
        // this row block's fragments in the stage buffers: B operand of this lane, ring code of this lane's column
        uint32_t sv = I.sv, sp = I.sp;
        const uint32_t spl = sp + (uint32_t)(nf - 1) * (4u * PS);  // last code of the block (look-ahead loads are clamped to it)

        // byte code -> shared address of the lane's A operand (m-tile 0): add the window base, wrap, add the lane's row
        auto conv = [&](uint32_t q) {
            const uint32_t v = q + baseB;
            return su32 + min(v, v - ringBr);  // v < 2 ringB: subtracts ringB iff v >= ringB (unsigned wrap otherwise)
        };
        auto flip = [&](double v) { return __hiloint2double(__double2hiint(v) ^ sgn, __double2loint(v)); };
        const bool work = active && nf > 0;
        // the operator fragments arrive ahead of the boxes: the first B operand and the ring addresses of the first two
        // fragments are fetched before the wait for the boxes
        // what-if 44 / 45 (no memory traffic): after the CTA's first tile a warp does not wait for any "full" barrier
        const bool nowait = (SBP_B3_EXP == 44 || SBP_B3_EXP == 45) && tile != (int)blockIdx.x;
        if (nowait) {
        } else if (kB3Merged) mbar_wait_a(bar32 + 8u * (O_FULL + st), phase);
        else mbar_wait_a(bar32 + 8u * (O_FFULL + fst), fphase);
        B3_TR(warp, trit, 1);
        double b0v = 0.0;
        uint32_t adr0 = su32, adr = su32;
        if (work) {
            b0v = flip(lds64<0>(sv));
            adr0 = conv(lds32<0>(sp));
            adr = conv(lds32<0>(min(sp + 4u * PS, spl)));  // address of fragment 1
        }
        if (!kB3Merged && !nowait) mbar_wait_a(bar32 + 8u * (O_FULL + st), phase);
        B3_TR(warp, trit, 2);
        if (PING) named_bar_sync(pp_me, 64);  // wait for the pipe token
        if (CW == 12 && p.sem) {  // block sem_n of this scheduler may multiply once sem_n - p.sem + 1... of its predecessors have finished
            const uint32_t a = smem_u32(&sem_fin[warp & 3]);
            int v;
            do {
                asm volatile("ld.volatile.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
            } while (v < sem_n - (p.sem - 1));
        }
        // ---- DMMA chain, two fragments per trip: operands one fragment ahead, ring codes two ahead ----
        if (work) {
            double a0[MT], a1[MT];
            // staggered chains: two free-running warps of a scheduler overlap their chains at random, and while both are in
            // their scalar phases the pipe idles; the gate keeps a warp from starting while its partner has just started
            constexpr bool GATE = SBP_B3_GATE > 0 && CW == kBsrCH;
            int got = 1;
            if (GATE && lane0) got = atomicCAS(&gate[warp & 3], -1, warp) == -1;  // first try, issued ahead of the operand loads
            lds_tiles<MT, MTB>(a0, adr0);
            if (GATE) {
                if (lane0)
                    while (!got) got = atomicCAS(&gate[warp & 3], -1, warp) == -1;
                __syncwarp();
            }
            B3_TR(warp, trit, 3);
            int rem = nf;
            const int stop = GATE ? nf - (((nf * SBP_B3_GATE) >> 3) & ~1) : nf;  // fragments left when the gate opens again
            auto trip = [&]() {
#if SBP_B3_EXP != 2 && SBP_B3_EXP != 41 && SBP_B3_EXP != 42
                lds_tiles<MT, MTB>(a1, adr);
#else
                a1[0] = __longlong_as_double(adr);
#endif
                const double b1v = flip(lds64<256>(sv));
                adr = conv(kNoCodes ? (uint32_t)rem * 32u : lds32<0>(min(sp + 8u * PS, spl)));
#pragma unroll
                for (int mt = 0; mt < MT; mt++) B3_DMMA(acc[mt][0], acc[mt][1], a0[mt], b0v);
                if (PRO && rem <= 3) {  // last trip: the pipe holds seven DMMAs of this warp -- the next item's integer work goes here
                    next_item();
                    nxt = prologue(r1, offB, svst, spst);
                    stepped = true;
                }
#if SBP_B3_EXP != 2 && SBP_B3_EXP != 41 && SBP_B3_EXP != 42
                lds_tiles<MT, MTB>(a0, adr);
#else
                a0[0] = __longlong_as_double(adr);
#endif
                // fragment 2 of this trip = fragment 0 of the next one (unused past the end)
                b0v = flip(lds64<512>(sv));
                adr = conv(kNoCodes ? (uint32_t)rem * 32u + 160u : lds32<0>(min(sp + 12u * PS, spl)));
#pragma unroll
                for (int mt = 0; mt < MT; mt++) B3_DMMA(acc[mt][0], acc[mt][1], a1[mt], b1v);
                sv += 512u, sp += 8u * PS, rem -= 2;
            };
            if (GATE) {
                while (rem > stop) trip();
                if (lane0) asm volatile("st.volatile.shared.s32 [%0], %1;\n" ::"r"(smem_u32(&gate[warp & 3])), "r"(-1) : "memory");
            }
            while (rem >= 2) trip();
            if (rem) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) B3_DMMA(acc[mt][0], acc[mt][1], a0[mt], b0v);
            }
        }

        if (PRO && !stepped) {  // blocks with fewer than two fragments, padding records
            next_item();
            nxt = prologue(r1, offB, svst, spst);
        }
        B3_TR(warp, trit, 4);
#if SBP_B3_RECPOS
        // the next block's record is requested now: its latency passes while the pipe drains and the results are staged
        const Rec r1 = load_rec(kBsrCH * pc + ps);
        advance();
#endif
        const int nb = min(Bir - b0c, TI);
        if (PING) named_bar_arrive(pp_next, 64);  // chain issued: the other warp of this scheduler may start
        if (CW == 12 && p.sem) {
            if (lane == 0) atomicAdd(&sem_fin[warp & 3], 1);
            sem_n += 3;
        }
        // blocks that carry SAT terms finish their accumulators now (the terms read the ring); for the others a deferred
        // epilogue also takes over the scaling by alpha
        const bool sat_block = (SAT == 1 && mask != 0) || (SAT == 2 && I.flags != 0);
        const bool scale_later = DEFER && scale_needed && !sat_block;
        if (XTc) {  // the helper has put this block's quads of `extra` into the staging buffer (every item signals)
            mbar_wait_a(wbar32 + 8u * O_XFULL, ext_phase);
            ext_phase ^= 1;
        }
        if (active) {
            if (scale_needed && !scale_later) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) acc[mt][0] *= p.alpha, acc[mt][1] *= p.alpha;
            }
            // element offset of node r of the lane's tensor row (m-tile 0) inside the ring: r lies in the chunk's window
            auto ring_elem = [&](int r) {
                const int q = r / BX;
                uint32_t o = baseB + (uint32_t)(q - I.lowbox) * BOXB + (uint32_t)(r - q * BX) * 8u;
                return min(o, o - ringB);
            };
            if (SAT == 1 && mask) {  // du_j += b_j (u_j - g_j); the chunk's own rows are inside its window
                const uint32_t o0 = ring_elem(r0w), o1 = ring_elem(r0w + 1);
                const double sbx = (FUSED && satfold) ? sb.x * p.ch : sb.x, sby = (FUSED && satfold) ? sb.y * p.ch : sb.y;
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    if (g_per_inst) {  // per-instance boundary data
                        const int il = 8 * mt + g;
                        const long long row = b0c + (il < nb ? il : 0);
                        if (mask & 1) gs0 = __ldg(p.g + (paired ? 2 * row + h0 : row) * p.g_stride + n0);
                        if (mask & 2) gs1 = __ldg(p.g + (paired ? 2 * row + h1 : row) * p.g_stride + n1);
                    }
                    if (satfold) {  // b_j u_j is part of the chain: -b_j g_j remains (times ch in a stage update)
                        if (mask & 1) acc[mt][0] = fma(-sbx, gs0, acc[mt][0]);
                        if (mask & 2) acc[mt][1] = fma(-sby, gs1, acc[mt][1]);
                    } else {
                        if (mask & 1) acc[mt][0] += sb.x * (*reinterpret_cast<const double*>(su_lane + o0 + mt * MTB) - gs0);
                        if (mask & 2) acc[mt][1] += sb.y * (*reinterpret_cast<const double*>(su_lane + o1 + mt * MTB) - gs1);
                    }
                }
            }
            if (SAT == 2 && I.flags != 0) {
                // 1D SATs at nodes 0 and Nh-1 of every instance (the record flags the block rows that carry one): lane l
                // evaluates the term for the warp's tensor rows l (and l + 32), the owning lane fetches it by shuffle
                const int flags = I.flags;
#pragma unroll 1
                for (int k = 0; k < 8; k++) {
                    const bool left = (flags >> k) & 1, right = (flags >> (8 + k)) & 1;
                    if (!(left | right)) continue;
                    const int r = (k < 4 ? I.rA : I.rB - 4) + k;
                    const int h = (paired && r >= Nh) ? 1 : 0;
                    // ring_elem is relative to the lane's own tensor row g; here lane l reads tensor row il
                    const uint32_t o = ring_elem(r);
                    double sv2[(8 * MT + 31) / 32];
#pragma unroll
                    for (int kk = 0; kk < (8 * MT + 31) / 32; kk++) {
                        const int il = lane + 32 * kk;  // tensor row within the tile
                        sv2[kk] = 0.0;
                        if (il < 8 * MT && il < nb) {
                            const long long row = (long long)b0c + il, inst = paired ? 2 * row + h : row;
                            const double u = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(s_u) + o +
                                                                              (size_t)il * (BX * 8));
                            double lb = sat_lb, rbv = sat_rb;
                            if (bc_per_inst) {  // per-instance boundary data
                                if (left) lb = __ldg(p.sat.bc + inst * p.sat.bc_stride);
                                if (right) rbv = __ldg(p.sat.bc + inst * p.sat.bc_stride + 1);
                            }
                            // (f - a u) / w as in sat1d_left/right, with the division replaced by the reciprocal weight
                            double v = 0.0;
                            if (left) v = sat_c0 * (lb - u);
                            if (right) v = fma(sat_cN, rbv - u, v);
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
            if (FUSED && r0w < Nr) {  // y <- ca * u + cb * extra + ch * y (u: the lane's own rows from the ring)
                const uint32_t o0 = ring_elem(r0w), o1 = ring_elem(r0w + 1);
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    double2 e = make_double2(0.0, 0.0);
                    if (with_extra) e = XTc ? *reinterpret_cast<const double2*>(sq + mt * 32) : ex[(FUSED && !HELP) ? mt : 0];
                    const double u0 = *reinterpret_cast<const double*>(su_lane + o0 + mt * MTB);
                    const double u1 = *reinterpret_cast<const double*>(su_lane + o1 + mt * MTB);
                    if (stagefold) {  // only the extra operand is left
                        if (XTRA) acc[mt][0] = fma(p.cb, e.x, acc[mt][0]), acc[mt][1] = fma(p.cb, e.y, acc[mt][1]);
                    } else {
                        acc[mt][0] = fma(p.ch, acc[mt][0], fma(p.ca, u0, p.cb * e.x));
                        acc[mt][1] = fma(p.ch, acc[mt][1], fma(p.ca, u1, p.cb * e.y));
                    }
                }
            }
        }
        __syncwarp();
        if (lane0) {
            mbar_arrive_a(bar32 + 8u * (O_EMPTY + stc));    // the ring slots behind this item's window may be refilled
            if (!kB3Merged) mbar_arrive_a(bar32 + 8u * (O_FEMPTY + fstc));  // and so may the fragment buffers
        }
        B3_TR(warp, trit, 5);

        if (SBP_B3_EXP == 4) {
            if (acc[0][0] == 1.2345e300) p.dU[0] = acc[MT - 1][1];
        } else if (DEFER) {
            if (pv_state & 1) emit(accs[DEFER ? hf ^ 1 : 0], pv_rowA, pv_rowB, pv_b0, (pv_state & 2) != 0);
            pv_rowA = I.rA, pv_rowB = I.rB, pv_b0 = b0c;
            pv_state = (active ? 1 : 0) | (scale_later ? 2 : 0) | (hf ? 4 : 0);
        } else if (HELP) {
            // ---- results to the warp's staging buffer; a helper warp stores them (every item signals, active or not) ----
            if (SBP_B3_EXP != 45) mbar_wait_a(wbar32 + 8u * O_OEMPTY, out_phase ^ 1);
            // a pair writes into both warps' buffers: the partner's previous contents must have left as well (same item, same parity);
            // after a pair item the helper has released both buffers at once, so the warp's own barrier covers the partner's
            if (pair_on && I.pair != 0 && !prev_pair) mbar_wait_a(wbar32 + 8u * O_OEMPTY + ((slot & 1) ? -8 : 8), out_phase ^ 1);
            if (pair_on) prev_pair = I.pair != 0;  // the helper has stored the previous contents (first item: passes)
            B3_TR(warp, trit, 6);
            out_phase ^= 1;
            if (active) {
                if (pair_on && I.pair != 0) {
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) *reinterpret_cast<double2*>(sqp + mt * 64) = make_double2(acc[mt][0], acc[mt][1]);
                } else if (MERGE8 && !XTc && I.rB == I.rA + 4) {  // [instance][8 rows]: lane (g, t) owns rows 2t, 2t + 1 of instance 8 mt + g
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) *reinterpret_cast<double2*>(sq8 + mt * 64) = make_double2(acc[mt][0], acc[mt][1]);
                } else {
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) *reinterpret_cast<double2*>(sq + mt * 32) = make_double2(acc[mt][0], acc[mt][1]);
                }
                if (!SBP_B3_HELPER_FENCE) fence_proxy_async();  // generic-proxy writes -> visible to the TMA (async proxy)
            }
            __syncwarp();  // orders the lanes' stores before lane 0's arrive (release)
            if (lane0) mbar_arrive_a(wbar32 + 8u * O_OFULL);
            B3_TR(warp, trit, 7);
        } else if (OUT == 2) {
            // ---- results through TMA: lane (g,t) deposits rows r0w, r0w+1 of instance 8mt + g in the warp's staging buffer;
            //      one lane stores quad A and quad B as {4 rows x 8*MT instances} boxes ----
            if (active) emit(acc, I.rA, I.rB, b0c, false);
        } else if (active && r0w < N) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const int il = 8 * mt + g;
                if (il >= nb) continue;
                double y0 = acc[mt][0], y1 = acc[mt][1];
                double* yp = p.dU + (long long)(b0c + il) * N + r0w;
                if (VEC) {  // N even: row r0w + 1 exists
                    if (p.beta != 0.0) {
                        const double2 o = *reinterpret_cast<const double2*>(yp);
                        y0 = fma(p.beta, o.x, y0);
                        y1 = fma(p.beta, o.y, y1);
                    }
                    stg_stream2(yp, y0, y1);
                } else {
                    if (p.beta != 0.0) y0 = fma(p.beta, yp[0], y0);
                    stg_stream1(yp, y0);
                    if (r0w + 1 < N) {
                        if (p.beta != 0.0) y1 = fma(p.beta, yp[1], y1);
                        stg_stream1(yp + 1, y1);
                    }
                }
            }
        }
        r0 = r1;
#if SBP_B3_TRACE
        trit++;
#endif
        skip = (myslot + CW) / kBsrCH - 1;
        myslot = (myslot + CW) % kBsrCH;
      } else {
        // no block of this item: observe its boxes (later windows reach back into them) and release it at once
        if (!kB3Merged) mbar_wait_a(bar32 + 8u * (O_FFULL + fst), fphase);
        mbar_wait_a(bar32 + 8u * (O_FULL + st), phase);
        if (lane0) {
            mbar_arrive_a(bar32 + 8u * (O_EMPTY + st));
            if (!kB3Merged) mbar_arrive_a(bar32 + 8u * (O_FEMPTY + fst));
        }
        skip--;
      }
        if (!PRO) next_item();
     }
    }
done:
    if (PING && (warp >> 2) == 0) named_bar_sync(pp_me, 64);  // absorb the last token
    if (DEFER && (pv_state & 1)) {  // the last block of this warp
        if (pv_state & 4) emit(accs[DEFER ? 1 : 0], pv_rowA, pv_rowB, pv_b0, (pv_state & 2) != 0);
        else emit(accs[0], pv_rowA, pv_rowB, pv_b0, (pv_state & 2) != 0);
    }
    if (OUT == 2 && !HELP && lane == 0) bulk_wait<0>();  // all output boxes of this warp have been written
}

static size_t bsr3_smem(const sbp_op* op, int MT, int D, int out, int cw) {
    const size_t ops = ((size_t)(kB3Merged ? D + 1 : kB3FragStages) * op->bsr_maxf * (32 * 8 + 4 * 4) + 127) & ~(size_t)127;
    // staging per consumer warp (TMA tensor stores only): 2 quads x instances x 4 rows
    const size_t staging = out == 2 ? (size_t)(cw == 12 ? 12 : kBsrCH) * 2 * 8 * MT * 4 * 8 : 0;
    return (size_t)op->bsr_nbox[D] * 8 * MT * kBsrBX * 8 + ops + staging;
}
static bool bsr3_fits(const sbp_ctx* ctx, const sbp_op* op, int MT, int D, int out, int cw) {
    return bsr3_smem(op, MT, D, out, cw) + 1024 <= (size_t)ctx->smem_optin;
}

template <int MT, int SAT, int OUT, int CW, bool HELP = false, bool FUSED = false, bool XTRA = false, bool GPI = false, bool PAIR = false>
static int32_t launch_bsr3_k(sbp_ctx* ctx, const BsrParams& p, size_t smem) {
    auto kern = k_bsr3<MT, SAT, OUT, CW, HELP, FUSED, XTRA, GPI, PAIR>;
    constexpr int TI = 8 * MT;
    constexpr int THREADS = (CW + 2 + (HELP ? 4 : 0)) * 32;
    CUtensorMap tmap, tmap_out;
    int32_t rc = make_tmap(&tmap, p.U, p.N, p.B, kBsrBX, TI);
    if (rc) return rc;
    rc = OUT == 2 ? make_tmap(&tmap_out, p.dU, p.N, p.B, 4, TI) : (tmap_out = tmap, SBP_OK);
    if (rc) return rc;
    CUtensorMap tmap_out8 = tmap_out;  // adjacent quads of a block as one box of 8 rows
    if (SBP_B3_MERGE8 && HELP && OUT == 2) rc = make_tmap(&tmap_out8, p.dU, p.N, p.B, 8, TI);
    if (rc) return rc;
    CUtensorMap tmap_ex = tmap_out;  // fused stage update with an extra operand: its quads travel like the output's
    if (FUSED && HELP && p.extra && p.cb != 0.0) rc = make_tmap(&tmap_ex, p.extra, p.N, p.B, 4, TI);
    if (SBP_B3_EXP == 11 || SBP_B3_EXP == 12) rc = make_tmap(&tmap_ex, p.dU, p.N, p.B, SBP_B3_EXP == 11 ? 32 : 16, TI);
    if (rc) return rc;
    SBP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = (p.B + TI - 1) / TI;
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)ctx->sm_count));
    kern<<<grid, THREADS, smem, ctx->stream>>>(p, tmap, tmap_out, tmap_ex, tmap_out8);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

template <int MT>
static int32_t launch_bsr3_mt(sbp_ctx* ctx, const BsrParams& p, size_t smem, int sat, int out, int cw) {
#define SBP_BSR3_OUT(S)                                                                                    \
    switch (out) {                                                                                         \
        case 2: return cw == 12 ? launch_bsr3_k<MT, S, 2, 12>(ctx, p, smem)                                \
                          : cw == 9 ? launch_bsr3_k<MT, S, 2, kBsrCH, true>(ctx, p, smem)                  \
                                    : launch_bsr3_k<MT, S, 2, kBsrCH>(ctx, p, smem);                       \
        case 1: return cw == 12 ? launch_bsr3_k<MT, S, 1, 12>(ctx, p, smem) : launch_bsr3_k<MT, S, 1, kBsrCH>(ctx, p, smem); \
        default: return launch_bsr3_k<MT, S, 0, kBsrCH>(ctx, p, smem);                                     \
    }
    if (p.pair8 && out == 2 && cw == 9 && sat != 2 && p.g_stride == 0) {  // x-neighbouring blocks stored as {8 x instances} boxes
        const bool xtra = p.extra != nullptr && p.cb != 0.0;
        if (!p.fused && sat == 0) return launch_bsr3_k<MT, 0, 2, kBsrCH, true, false, false, false, true>(ctx, p, smem);
        if (!p.fused && sat == 1) return launch_bsr3_k<MT, 1, 2, kBsrCH, true, false, false, false, true>(ctx, p, smem);
        if (p.fused && sat == 1 && !xtra) return launch_bsr3_k<MT, 1, 2, kBsrCH, true, true, false, false, true>(ctx, p, smem);
    }
    if (p.fused) {  // stage updates: RHS launches with TMA output and store helpers (checked by launch_bsr3)
        const bool xtra = p.extra != nullptr && p.cb != 0.0;
        if (sat == 1 && p.g_stride != 0)
            return xtra ? launch_bsr3_k<MT, 1, 2, kBsrCH, true, true, true, true>(ctx, p, smem)
                        : launch_bsr3_k<MT, 1, 2, kBsrCH, true, true, false, true>(ctx, p, smem);
        if (sat == 1) return xtra ? launch_bsr3_k<MT, 1, 2, kBsrCH, true, true, true>(ctx, p, smem) : launch_bsr3_k<MT, 1, 2, kBsrCH, true, true>(ctx, p, smem);
        if (p.sat.bc_stride != 0)
            return xtra ? launch_bsr3_k<MT, 2, 2, kBsrCH, true, true, true, true>(ctx, p, smem)
                        : launch_bsr3_k<MT, 2, 2, kBsrCH, true, true, false, true>(ctx, p, smem);
        return xtra ? launch_bsr3_k<MT, 2, 2, kBsrCH, true, true, true>(ctx, p, smem) : launch_bsr3_k<MT, 2, 2, kBsrCH, true, true>(ctx, p, smem);
    }
    if (sat == 1 && out == 2 && cw == 9 && p.g_stride != 0) return launch_bsr3_k<MT, 1, 2, kBsrCH, true, false, false, true>(ctx, p, smem);
    if (sat == 2 && out == 2 && cw == 9 && p.sat.bc_stride != 0) return launch_bsr3_k<MT, 2, 2, kBsrCH, true, false, false, true>(ctx, p, smem);
    if (sat == 1) { SBP_BSR3_OUT(1) }
    if (sat == 2) { SBP_BSR3_OUT(2) }
    SBP_BSR3_OUT(0)
#undef SBP_BSR3_OUT
}

// image of a stage update: ch * (A - B) - ca * I in fragment form (the chain, run with alpha = -1, then yields ca u - ch (A - B) u)
__global__ void k_stage_image(double* __restrict__ out, const double* __restrict__ in, const unsigned char* __restrict__ mask,
                              long long n, double ch, double ca) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ch * in[i] - ((mask && mask[i]) ? ca : 0.0);
}

// p arrives filled in by launch_bsr (operator, batch view, SAT data); returns false when this kernel does not apply
bool launch_bsr3(sbp_ctx* ctx, const sbp_op* op, BsrParams& p, int satmode, int out, int32_t* rc) {
    // option bsr_v3 (SBP_BSR_V3): 0 = k_bsr, 1 = this kernel for large batches, 2 = always
    const int forced_mt = ctx->tune[T_BSR_MT], forced_d = ctx->tune[T_BSR_DEPTH];
    // Default rule (measured, profiles/r1k_bsr3_notes.txt section 5): with the store-helper warps this kernel beats k_bsr on
    // every RHS launch (2D or 1D SATs fused) and on instance-pair views (odd N), by 2-18 %; k_bsr stays ahead on plain
    // even-N applies.  option bsr_v3 (SBP_BSR_V3): -1 = that rule, 0 = always k_bsr,
    // 1 = this kernel for every large batch, 2 = always this kernel.
    const int enabled = ctx->tune[T_BSR_V3];
    if (!enabled || kBsrGroups != 1 || op->d_bsr_pos3 == nullptr) return false;
    // (r2: also plain applies of operators whose tile row is much longer than the ring -- 2D clouds: config 3's A without the
    // SAT terms 0.322 ms on k_bsr, 0.280 here)
    const bool wide = p.nbt > 2 * op->bsr_nbox[1];
    if (enabled < 0 && !(out == 2 && (satmode != 0 || op->bsr_N != op->N || wide))) return false;
    const int forced_cw = ctx->tune[T_BSR_CW], help = ctx->tune[T_BSR_HELP];
    // cw: 8 consumer warps | 12 (SBP_BSR_CW=12) | 9 = 8 consumer warps + store helper warps (SBP_BSR_HELP=1, TMA output)
    const int cw = (forced_cw == 12 && out != 0 && !p.fused && kBsrCH == 8) ? 12 : (help && out == 2) ? 9 : 8;
    // fused stage updates: helper variant of an RHS launch (the extra operand travels through the staging buffers: the
    // helpers fetch its quads by TMA as soon as the previous block's results have been read out)
    // (measured: 2D RHS N = 1296 0.362 ms per stage with the extra operand against 0.414 on k_bsr; 1D RHS N = 101 0.279 against
    // 0.240 -> those stay on k_bsr)
    if (p.fused && (cw != 9 || satmode == 0 || (satmode == 2 && p.cb != 0.0 && enabled != 2))) return false;
    // the helper kernels with 2D SATs are compiled for the folded image A - diag(b) (needs alpha * scale = -1: the RHS entry points)
    if (SBP_B3_SATFOLD && SBP_B3_EXP == 0 && satmode == 1 && cw == 9 && (!p.val_sat || p.alpha != -1.0)) return false;
    const long long B = p.B;
    // the largest tile with one chunk of lookahead that still gives every SM work, then the deepest ring that fits
    int MT = 8;
    while (MT > 4 && !bsr3_fits(ctx, op, MT, 1, out, cw)) MT--;
    while (MT > 4 && (B + 8 * MT - 1) / (8 * MT) < ctx->sm_count) MT--;
    if (forced_mt >= 4 && forced_mt <= 8) MT = forced_mt;
    if (!bsr3_fits(ctx, op, MT, 1, out, cw)) return false;
    if (enabled != 2 && (B + 8 * MT - 1) / (8 * MT) < ctx->sm_count) return false;  // small batches: smaller tiles of k_bsr
    int D = 1;
    for (int d = 2; d < kB3Stages; d++)
        if (d <= forced_d && bsr3_fits(ctx, op, MT, d, out, cw)) D = d;
    p.nbox = op->bsr_nbox[D];
    p.nst = D + 1;
    p.item = reinterpret_cast<const int4*>(op->d_bsr_item) + (size_t)D * p.nch * kBsrCH * 2;
    p.pos = op->d_bsr_pos3 + (size_t)(MT - 1) * (size_t)std::max<int64_t>(op->bsr_frags, 1) * 4;
    const size_t smem = bsr3_smem(op, MT, D, out, cw);
    p.skew = ctx->tune[T_BSR_SKEW];
    p.sem = ctx->tune[T_BSR_SEM];
    // x-neighbouring blocks as {8 x instances} boxes: pays where the stores weigh (regular 2D grids: 4.9 fragments per block,
    // 0.251 -> 0.238 ms), not on the fragment-heavy jittered cloud of config 3 (9 per block: 0.2826 -> 0.2847); option bsr_pair8
    p.pair8 = ctx->tune[T_BSR_PAIR8] >= 0 ? ctx->tune[T_BSR_PAIR8]
                                          : (op->bsr_pairs * 4 >= op->bsr_nrb && op->bsr_frags <= 7 * (int64_t)op->bsr_nrb);
    if (SBP_B3_SATFOLD && SBP_B3_EXP == 0 && p.fused && satmode == 2 && cw == 9) {
        // 1D stage update: image ch alpha D + ca I (diagonal mask of the operator uploaded at first use), run with alpha = +1
        sbp_op* mop = const_cast<sbp_op*>(op);
        const long long nv = (long long)op->bsr_frags * 32;
        if (!mop->d_bsr_diagmask) {
            std::vector<unsigned char> mask((size_t)nv, 0);
            for (int k : op->bsr_diag)
                if (k >= 0) mask[k] = 1;
            if (cudaMalloc(&mop->d_bsr_diagmask, (size_t)nv) != cudaSuccess ||
                cudaMemcpy(mop->d_bsr_diagmask, mask.data(), (size_t)nv, cudaMemcpyHostToDevice) != cudaSuccess) {
                *rc = SBP_ENOMEM;
                return true;
            }
        }
        if (!mop->d_bsr_val_stage && cudaMalloc(&mop->d_bsr_val_stage, (size_t)nv * 8) != cudaSuccess) {
            *rc = SBP_ENOMEM;
            return true;
        }
        k_stage_image<<<(unsigned)((nv + 255) / 256), 256, 0, ctx->stream>>>(mop->d_bsr_val_stage, p.val, mop->d_bsr_diagmask, nv,
                                                                             p.ch * p.alpha, -p.ca);
        ctx->launches++;
        p.val = mop->d_bsr_val_stage;
        p.alpha = 1.0;
    }
    // alpha other than +-1: the launch runs on a scaled copy of the image (3 us) instead of multiplying 14 accumulators per block
    if (p.alpha != 1.0 && p.alpha != -1.0 && !(SBP_B3_SATFOLD && satmode == 1 && cw == 9)) {
        sbp_op* mop = const_cast<sbp_op*>(op);  // scratch of the operator (launches of a context are stream-ordered)
        const long long nv = (long long)op->bsr_frags * 32;
        if (!mop->d_bsr_val_stage && cudaMalloc(&mop->d_bsr_val_stage, (size_t)nv * 8) != cudaSuccess) {
            *rc = SBP_ENOMEM;
            return true;
        }
        k_stage_image<<<(unsigned)((nv + 255) / 256), 256, 0, ctx->stream>>>(mop->d_bsr_val_stage, p.val, nullptr, nv, p.alpha, 0.0);
        ctx->launches++;
        p.val = mop->d_bsr_val_stage;
        p.alpha = 1.0;
    }
    // 2D SATs with the u-term folded into the operator image (SBP_B3_SATFOLD; needs alpha * scale = -1: the RHS entry points)
    if (SBP_B3_SATFOLD && SBP_B3_EXP == 0 && satmode == 1 && cw == 9) {
        p.satfold = 1;
        p.val = p.val_sat;
        if (p.fused) {
            k_stage_image<<<(unsigned)((p.nvals + 255) / 256), 256, 0, ctx->stream>>>(p.val_stage, p.val_sat, p.diagmask, p.nvals,
                                                                                       p.ch, p.ca);
            ctx->launches++;
            p.val = p.val_stage;
        }
    }
    // L2 prefetch distance of the box producer: 5 boxes ahead when a tile row is much longer than the ring (2D clouds: config 3
    // 0.312 -> 0.291 ms, N = 900 / 1936 -5 / -4 %); banded 1D operators, whose ring holds most of the row, lose 6-13 % with any
    // prefetch (profiles/r2_bsr_notes.txt)
    p.pf = std::min(ctx->tune[T_BSR_PF] >= 0 ? ctx->tune[T_BSR_PF] : (p.nbt > 2 * p.nbox ? 5 : 0), p.nbt - 1);
    const bool verbose = ctx->tune[T_BSR_VERBOSE] != 0;
    if (verbose)
        fprintf(stderr, "k_bsr3: N=%d B=%lld nrb=%d nch=%d frags=%lld maxnf=%d MT=%d depth=%d nbox=%d smem=%zu sat=%d out=%d cw=%d satfold=%d fused=%d\n",
                p.N, B, p.nrb, p.nch, (long long)op->bsr_frags, op->bsr_maxnf, MT, D, p.nbox, smem, satmode, out, cw, p.satfold, p.fused);
    switch (MT) {
        case 8: *rc = launch_bsr3_mt<8>(ctx, p, smem, satmode, out, cw); break;
        case 7: *rc = launch_bsr3_mt<7>(ctx, p, smem, satmode, out, cw); break;
        case 6: *rc = launch_bsr3_mt<6>(ctx, p, smem, satmode, out, cw); break;
        case 5: *rc = launch_bsr3_mt<5>(ctx, p, smem, satmode, out, cw); break;
        default: *rc = launch_bsr3_mt<4>(ctx, p, smem, satmode, out, cw); break;
    }
    return true;
}

}  // namespace sbp

#if SBP_B3_TRACE
extern "C" int32_t sbp_debug_set_trace(void* dptr) {
    return cudaMemcpyToSymbol(sbp::g_b3_trace, &dptr, sizeof(void*)) == cudaSuccess ? 0 : -3;
}
#endif
