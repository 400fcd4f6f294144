// bsr.cuh -- kernel parameters and TMA tensor-copy wrappers shared by the block-sparse DMMA kernels (bsr.cu, bsr3.cu)
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "ptx.cuh"

namespace sbp {

struct BsrParams {
    const double* __restrict__ U;
    double* __restrict__ dU;
    const double* __restrict__ val;   // [fragment][32]
    const int* __restrict__ pos;      // [fragment][4]  (box - boxlo[chunk]) << 8 | node % kBsrBX
    const int4* __restrict__ item;    // [nch * kBsrCH][2] consumer records for this launch's pipeline depth (api.cu build_bsr)
    const int4* __restrict__ rbmeta;  // [nrb] {first fragment relative to the chunk, fragments, first row of quad A, of quad B}
    const int4* __restrict__ chmeta;  // [nch] {first fragment, end fragment, boxlo, boxhi}: boxes [0, boxhi) of the tile are
                                      // resident once the chunk has been staged; no later chunk touches a box below boxlo
    int N, nrb, nch, nbox, nbt, maxf, nst;  // nbox ring slots, nbt boxes per tile, nst pipeline stages
    int Nh;                           // nodes per instance: N, or N/2 when a tensor row holds an instance PAIR (odd Nh)
    long long B;                      // tensor rows (instances, or instance pairs)
    double alpha, beta;
    // fused Runge-Kutta stage update (sbp_rhs_advection*_stage): dU <- ca * U + cb * extra + ch * (RHS); U's own rows come from
    // the ring, extra (may alias dU, not U) is read by the lane that writes the same entries
    const double* __restrict__ extra = nullptr;
    double ca = 0.0, cb = 0.0, ch = 1.0;
    int fused = 0;
    int skew = 0;                     // bsr3.cu: start-up delay (cycles) of the second consumer warp of every scheduler
    int pf = 0;                       // bsr3.cu: L2 prefetch distance of the box producer, in boxes of its stream (0: none)
    // SAT u-term folded into the operator (k_bsr3 with 2D SATs): val_sat = fragments of A - diag(b); stage updates run on
    // val_stage = ch * val_sat - ca * (diagonal mask), written by a small kernel in front of the launch
    const double* val_sat = nullptr;
    double* val_stage = nullptr;
    const unsigned char* diagmask = nullptr;
    long long nvals = 0;
    int satfold = 0;
    int pair8 = 0;                    // 1: pair-flagged blocks are stored as {8 x instances} boxes; the two warps of a scheduler take neighbouring slots                  // 1: the image in `val` contains the fold (set by launch_bsr3)
    int sem = 0;                      // bsr3.cu, 12 consumer warps: at most `sem` of the three warps of a scheduler multiply at a time (0: off)
    // fused 2D SAT: du_j += b_j (u_j - g_j) on inflow rows
    const double* __restrict__ rb_bdiag;  // [nrb][8] b_j of the block's rows in block order, 0 on non-inflow rows
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

// TMA 2-D tensor store: box {4 rows, instances} from shared memory -> dU (bulk async-group completion); rows beyond N
// and instances beyond B are clipped by the hardware
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, int x, int y, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(tmap), "r"(x),
                 "r"(y), "r"(smem_u32(smem_src))
                 : "memory");
}

// 2-D tensor map over a row-major [B][N] array of doubles with boxes of box_nodes x box_inst (bsr.cu)
int32_t make_tmap(CUtensorMap* tm, const double* base, int N, long long B, int box_nodes, int box_inst);

}  // namespace sbp
