// common.cuh -- internal structs and error plumbing of libsbp_b200.so
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/sbp_b200.h"

namespace sbp {

void set_error(const char* fmt, ...);

#define SBP_CUDA(call)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            sbp::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return (e_ == cudaErrorMemoryAllocation) ? SBP_ENOMEM : SBP_ECUDA;                    \
        }                                                                                         \
    } while (0)

#define SBP_REQUIRE(cond, code, ...)     \
    do {                                 \
        if (!(cond)) {                   \
            sbp::set_error(__VA_ARGS__); \
            return (code);               \
        }                                \
    } while (0)

// block-sparse kernel (bsr.cu) geometry, shared by the host-side image builder (api.cu) and the kernel
#ifndef SBP_BSR_BX
#define SBP_BSR_BX 20
#endif
#ifndef SBP_BSR_CH
#define SBP_BSR_CH 8
#endif
#ifndef SBP_BSR_PING
#define SBP_BSR_PING 0  // 1: consumer warps of one scheduler take turns on the FP64 tensor pipe (see bsr.cu; measured: no gain)
#endif
#ifndef SBP_BSR_G
#define SBP_BSR_G 1
#endif
constexpr int kBsrBX = SBP_BSR_BX;     // nodes per TMA box: 4 (mod 16) doubles => conflict-free A-fragment gathers
constexpr int kBsrCH = SBP_BSR_CH;     // row blocks per chunk
constexpr int kBsrGroups = SBP_BSR_G;  // consumer groups, alternating chunks
constexpr int kBsrDepths = 8;          // pipeline depths (chunks of lookahead) the host prepares ring sizes / records for
static_assert(kBsrBX % 16 == 4 || kBsrBX % 16 == 12, "box rows must be 4 or 12 (mod 16) doubles: instance g then starts at bank 8g (or -8g) mod 32");
constexpr int kMaxSmallN = 128;  // smem-resident DMMA path handles N <= 128 (16 n-tiles of 8 rows)

// Tuning / debugging knobs.  Each is read from its environment variable ONCE, when a context is created, and can be changed
// per context with sbp_ctx_set_option (the tests and the sweep tools do); nothing on a launch path calls getenv.
enum Tune {
    T_SPARSE_LOGR,     // SBP_SPARSE_LOGR     -1    sparse.cu: force rows per lane = 1 << value
    T_BSR_NO_PAIRING,  // SBP_BSR_NO_PAIRING   0    build_bsr: pair each quad with its successor only
    T_SPARSE_KERNEL,   // SBP_SPARSE_KERNEL    1    kernel of a sparse operator chosen at upload: 1 / "bsr" block-sparse DMMA, 0 / "sell"
    T_SYM_VARIANT,     // SBP_SYM_VARIANT      1    even/odd kernel: bit0 prefetch, bit1 one m-tile per warp, bit2 256-thread CTAs
    T_FUSED_STAGE,     // SBP_FUSED_STAGE      1    0: stage updates as RHS + sbp_lincomb
    T_SOLVE_SMALL,     // SBP_SOLVE_SMALL     -1    single-launch solve: -1 by batch size, 0 never, 1 whenever applicable
    T_BSR_OUT,         // SBP_BSR_OUT          2    2 TMA tensor stores, 1 plain 16-byte stores
    T_BSR_MT,          // SBP_BSR_MT           0    force the m-tiles per warp (tile = 8 MT instances)
    T_BSR_DEPTH,       // SBP_BSR_DEPTH       99    cap of the box pipeline depth
    T_BSR_MH,          // SBP_BSR_MH           0    2: two warps per row block (k_bsr)
    T_BSR_V3,          // SBP_BSR_V3          -1    -1 default rule, 0 always k_bsr, 1 k_bsr3 for large batches, 2 always k_bsr3
    T_BSR_CW,          // SBP_BSR_CW           8    consumer warps of k_bsr3 (8 or 12)
    T_BSR_HELP,        // SBP_BSR_HELP         1    store-helper warps
    T_BSR_SKEW,        // SBP_BSR_SKEW         0    start-up delay (cycles) of every second consumer warp
    T_BSR_SEM,         // SBP_BSR_SEM          0    12 consumer warps: at most this many of a scheduler's three multiply at a time
    T_BSR_PF,          // SBP_BSR_PF          -1    L2 prefetch distance of the box producer in boxes (-1: by the shape of the tile)
    T_BSR_VERBOSE,     // SBP_BSR_VERBOSE      0    print the launch geometry
    T_GEMM_VARIANT,    // SBP_GEMM_VARIANT    -2    < 0: TMA pipeline, >= 0: one of the cp.async configurations
    T_DS_V4,           // SBP_DS_V4            1    256-bit small-dense kernel variant (-1: off)
    T_SPARSE_DB,       // SBP_SPARSE_DB        0    double buffering of the SELL kernel
    T_BSR_PAIR8,       // SBP_BSR_PAIR8       -1    x-neighbouring blocks stored as {8 x instances} boxes: -1 by the operator, 0 off, 1 on
    T_COUNT
};
struct TuneSpec {
    const char* name;  // option name for sbp_ctx_set_option ("bsr_v3", ...)
    const char* env;
    int def;
};
extern const TuneSpec kTuneSpec[T_COUNT];

}  // namespace sbp

struct sbp_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // stream launches go to (own_stream or caller-provided)
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams for the host-buffer pipeline
    int sm_count = 0;
    int64_t l2_bytes = 0;
    int64_t smem_optin = 0;
    int64_t global_bytes = 0;
    int64_t launches = 0;
    double* scratch = nullptr;  // per-CTA partial sums for deterministic batch reductions
    size_t scratch_doubles = 0;
    double* stage_buf = nullptr;  // RHS of the un-fused stage updates (sbp_rhs_advection*_stage fallback)
    size_t stage_doubles = 0;
    // host pipeline staging
    void* pipe_dev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [slot][in/out]
    size_t pipe_bytes = 0;
    cudaEvent_t pipe_ev[2][3] = {};
    void* nccl_comm = nullptr;
    int tune[sbp::T_COUNT] = {};  // tuning knobs (sbp::Tune), environment defaults taken at creation
    // grow-only device buffers of the *_host / solve entry points: [0] boundary data, [1] Runge-Kutta work vectors, [2] state
    // [3] step sizes of the single-launch solve
    void* aux[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t aux_bytes[4] = {0, 0, 0, 0};
};

struct sbp_op {
    int kind = 0;
    int N = 0;
    int G = 1;
    int64_t nnz = 0;
    int variant = 0;
    sbp_ctx* ctx = nullptr;
    double scale = 1.0;
    // dense / table
    double* d_D = nullptr;     // G matrices, ROW-major N x N each (row i contiguous in k), scale NOT folded
    double* d_frag = nullptr;  // DMMA B-fragment order, G * (KS*NT*32) doubles (small N only)
    double* d_wfrag = nullptr; // weights in A-fragment k order, G * (KS*4) doubles
    double* d_w = nullptr;     // G*N weights (or nullptr)
    int NT = 0;                // n-tiles of 8 rows (small path)
    bool sym = false;          // centro-antisymmetric (even/odd split available)
    double* d_frag_sym = nullptr;  // [E | O] half-operator fragments (dense_small_sym.cu)
    double* d_wsym = nullptr;      // weights [w(j) | w(N-1-j)] in fragment k order
    int layout = 2;                // fragment layout of d_frag (dense_small.cu): 2 paired-k (N % 8 == 0), 1 natural
    int band = 0;                  // > 0: fragments with |k-pair - n-tile| > band are structurally zero (banded D)
    double zero_frac = 0.0;        // fraction of structurally zero fragments
    double* d_frag4 = nullptr;     // fragment image in "layout 4" (256-bit accesses, dense_small_v4.cu), N % 16 == 0
    double* d_wfrag4 = nullptr;
    // sparse (SELL-32 slices)
    int* d_rowptr = nullptr;
    int* d_col = nullptr;
    double* d_val = nullptr;
    int n_slices = 0;
    int logR = 0;                // rows per lane = 1 << logR (row-blocked SELL, sparse.cu)
    int* d_slice_ptr = nullptr;  // n_slices+1 offsets (in units of 32-wide columns)
    int* d_sell_col = nullptr;
    double* d_sell_val = nullptr;
    int max_row_nnz = 0;
    // plain CSR copy (row-wise consumers: the single-launch solve of small batches, solve_small.cu)
    int* d_csr_rowptr = nullptr;
    int* d_csr_col = nullptr;
    double* d_csr_val = nullptr;
    int csr_max_row = 0;  // most non-zeros of a row
    // sparse, block-sparse DMMA image (bsr.cu): row blocks of 8 rows, fragments of 4 gathered columns
    double* d_bsr_val = nullptr;  // [fragment][32] B-operand fragments (row g, column t)
    int* d_bsr_pos = nullptr;     // [fragment][4] (box - boxlo[chunk]) << 8 | column % kBsrBX
    int* d_bsr_rbmeta = nullptr;  // [nrb][4] {first fragment relative to its chunk, fragments, first row of quad A, of quad B}
    int* d_bsr_item = nullptr;    // [kBsrDepths][nch * kBsrCH][8] per-row-block record of the consumer warps (see build_bsr)
    int* d_bsr_pos3 = nullptr;    // [8 tile sizes MT][fragment][4] BYTE offsets (box - boxlo[chunk]) * box bytes(MT) + 8 * column (bsr3.cu)
    int* d_bsr_chmeta = nullptr;  // [nch][4] {first fragment, end fragment, lowest box still needed, boxes resident}
    int bsr_nrb = 0;
    int bsr_N = 0;                // row length the block-sparse image works on: N, or 2N for odd N (instance pairs, see build_bsr)
    std::vector<int> bsr_rows;    // host copy: first rows of the two quads of every row block (2 * nrb)
    int bsr_pairs = 0;            // row blocks flagged as halves of x-neighbouring pairs (bsr3.cu, PAIR8)
    std::vector<int> bsr_diag;    // host copy: index into d_bsr_val of the diagonal entry of every block row (8 * nrb, -1: row >= N)
    int bsr_nbox[sbp::kBsrDepths] = {};  // ring size (boxes) that allows d chunks of lookahead, d = 0..kBsrDepths-1
    int bsr_maxf = 0;             // most fragments of a chunk
    int bsr_maxnf = 0;            // most fragments of a row block
    int64_t bsr_frags = 0;
    // advection bundles
    const sbp_op* inner = nullptr;
    double* d_b = nullptr;
    int* d_mode = nullptr;
    int Nb = 0;
    int* d_bnd_node = nullptr;
    double* d_bnd_tau = nullptr;
    double* d_a = nullptr;
    double* d_q_consts = nullptr;  // per-node constants for the analysis quantities
    double* d_rb_bdiag = nullptr;  // 2D bundle over a block-sparse operator: b_j of every block row (block order), 0 = no SAT
    // the same bundle with the SAT's u-term folded into the operator: fragments of A - diag(b) (du = -(A - B) u - B g), a scratch
    // image for stage updates (ch (A - B) - ca I, rebuilt per launch) and the 0/1 mask of the diagonal entries
    double* d_bsr_val_sat = nullptr;
    double* d_bsr_val_stage = nullptr;
    unsigned char* d_bsr_diagmask = nullptr;
    double sat_a0 = 1.0, sat_aN = 1.0, sat_w0 = 1.0, sat_wN = 1.0;  // 1D bundle: end speeds and end weights
};

namespace sbp {
// Fused SATs of the upstream 1D semidiscretization: applied by the thread that owns row 0 / row N-1.
//   du[0]   += (fl - a0*u0)/w0,   fl = godunov(bc_left, u0, a0)  = a0 > 0 ? a0*bc_left : a0*u0
//   du[N-1] -= (fr - aN*uN)/wN,   fr = godunov(uN, bc_right, aN) = aN > 0 ? aN*uN     : aN*bc_right
struct Sat1D {
    const double* bc = nullptr;  // [left, right] per instance (stride bc_stride doubles) or shared (stride 0)
    long long bc_stride = 0;
    double a0 = 1.0, aN = 1.0, w0 = 1.0, wN = 1.0;
    int enabled = 0;
};
#ifdef __CUDACC__
// t0 = a0*u0 (as rounded by the reference's a .* u), returns the term ADDED to du[0]
__device__ __forceinline__ double sat1d_left(const Sat1D& s, long long inst, double t0) {
    const double lb = s.bc[inst * s.bc_stride];
    const double fl = s.a0 > 0.0 ? s.a0 * lb : t0;
    return (fl - t0) / s.w0;
}
__device__ __forceinline__ double sat1d_right(const Sat1D& s, long long inst, double tN) {
    const double rb = s.bc[inst * s.bc_stride + 1];
    const double fr = s.aN > 0.0 ? tN : s.aN * rb;
    return -((fr - tN) / s.wN);
}
// same with the boundary value already in a register
__device__ __forceinline__ double sat1d_left_v(const Sat1D& s, double lb, double t0) {
    const double fl = s.a0 > 0.0 ? s.a0 * lb : t0;
    return (fl - t0) / s.w0;
}
__device__ __forceinline__ double sat1d_right_v(const Sat1D& s, double rb, double tN) {
    const double fr = s.aN > 0.0 ? tN : s.aN * rb;
    return -((fr - tN) / s.wN);
}
#endif
// kernel-family launchers (each in its own .cu)
int32_t launch_dense_small(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                           double beta, bool periodic_table, double* energy, double* energy_sum);
int32_t launch_dense_small_sym(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                               double beta, double* energy, double* energy_sum, int variant,
                               const Sat1D* sat = nullptr);
int32_t launch_dense_small_v4(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                              double beta, double* energy, double* energy_sum, int variant);
int32_t launch_table_generic(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B,
                             const int32_t* op_index, double alpha, double beta, double* energy, double* energy_sum);
int32_t launch_dense_gemm(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                          double beta);
int32_t launch_sparse(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                      double beta, const sbp_op* adv, const double* g, int64_t g_stride, double* quantities);
bool bsr_available(const sbp_ctx* ctx, const sbp_op* op);
bool bsr_can_run(const sbp_op* op, const double* U);
struct BsrParams;
// bsr3.cu: lean consumer loop (operator fragments prefetched from L2 into registers); false = not applicable, use k_bsr
bool launch_bsr3(sbp_ctx* ctx, const sbp_op* op, BsrParams& p, int satmode, int out, int32_t* rc);
// Runge-Kutta stage update fused into an RHS kernel: out <- ca * U + cb * extra + ch * RHS(U)  (extra may be null: cb ignored)
struct StageUpdate {
    double ca, cb, ch;
    const double* extra;
};
int32_t launch_bsr(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha, double beta,
                   const sbp_op* adv, const double* g, int64_t g_stride, const Sat1D* sat,
                   const StageUpdate* su = nullptr);
int32_t launch_integrate(sbp_ctx* ctx, const double* d_w, int N, int G, const double* U, const double* V, int64_t B,
                         int mode, double* out, double* out_sum);
int32_t launch_scale(sbp_ctx* ctx, const double* d_w, int N, double* U, int64_t B, double factor, int inverse);
int32_t launch_axpbypcz(sbp_ctx* ctx, double* Y, const double* X, const double* Z, int64_t n, double a, double b,
                        double c);
int32_t launch_lincomb(sbp_ctx* ctx, double* Y, const double* X, const double* Z, const double* W, int64_t n, double a,
                       double b, double c, double d);
int32_t launch_sum_partials(sbp_ctx* ctx, const double* partials, int n, double* out);
int32_t launch_adv2d_epilogue(sbp_ctx* ctx, const sbp_op* adv, double* dU, const double* U, int64_t B,
                              const double* g, int64_t g_stride, double* quantities, bool apply_sat);
int32_t launch_adv1d(sbp_ctx* ctx, const sbp_op* adv, double* dU, const double* U, int64_t B, const double* bc,
                     int64_t bc_stride, double* quantities);
int32_t ensure_scratch(sbp_ctx* ctx, size_t doubles);
int32_t ensure_aux(sbp_ctx* ctx, int slot, size_t bytes, void** out);  // grow-only device buffer sbp_ctx::aux[slot]
// solve_small.cu: a whole SSPRK53 solve of a small batch in one launch (one CTA per instance, state in shared memory)
bool solve_small_applies(const sbp_ctx* ctx, const sbp_op* op, int64_t B);
int32_t launch_solve_small(sbp_ctx* ctx, const sbp_op* op, double* U, int64_t B, const double* bc_table,
                           int64_t bc_row_stride, int64_t nsteps, const double* d_hsteps);
}  // namespace sbp
