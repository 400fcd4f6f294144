// api.cu -- extern "C" entry points of libsbp_b200.so (see include/sbp_b200.h for the contract and the
// reference file:line each entry point replaces).
#include <algorithm>
#include <iterator>
#include <cmath>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>
#include "common.cuh"

namespace sbp {

static thread_local std::string g_last_error;

const TuneSpec kTuneSpec[T_COUNT] = {
    {"sparse_logr", "SBP_SPARSE_LOGR", -1}, {"bsr_no_pairing", "SBP_BSR_NO_PAIRING", 0}, {"sparse_kernel", "SBP_SPARSE_KERNEL", 1},
    {"sym_variant", "SBP_SYM_VARIANT", 1},  {"fused_stage", "SBP_FUSED_STAGE", 1},       {"solve_small", "SBP_SOLVE_SMALL", -1},
    {"bsr_out", "SBP_BSR_OUT", 2},          {"bsr_mt", "SBP_BSR_MT", 0},                 {"bsr_depth", "SBP_BSR_DEPTH", 99},
    {"bsr_mh", "SBP_BSR_MH", 0},            {"bsr_v3", "SBP_BSR_V3", -1},                {"bsr_cw", "SBP_BSR_CW", 8},
    {"bsr_help", "SBP_BSR_HELP", 1},        {"bsr_skew", "SBP_BSR_SKEW", 0},             {"bsr_sem", "SBP_BSR_SEM", 0},
    {"bsr_pf", "SBP_BSR_PF", -1},           {"bsr_verbose", "SBP_BSR_VERBOSE", 0},       {"gemm_variant", "SBP_GEMM_VARIANT", -2},
    {"ds_v4", "SBP_DS_V4", 1},              {"sparse_db", "SBP_SPARSE_DB", 0},           {"bsr_pair8", "SBP_BSR_PAIR8", -1}};

static void tune_from_env(sbp_ctx* ctx) {
    for (int k = 0; k < T_COUNT; k++) {
        const char* e = getenv(kTuneSpec[k].env);
        int v = kTuneSpec[k].def;
        if (e) {
            if (k == T_SPARSE_KERNEL && (strcmp(e, "sell") == 0 || strcmp(e, "bsr") == 0)) v = strcmp(e, "bsr") == 0;
            else if (k == T_BSR_NO_PAIRING || k == T_BSR_VERBOSE) v = (*e && strcmp(e, "0") != 0) ? 1 : 0;
            else v = atoi(e);
        }
        ctx->tune[k] = v;
    }
}

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int32_t ensure_scratch(sbp_ctx* ctx, size_t doubles) {
    if (ctx->scratch_doubles >= doubles) return SBP_OK;
    // grown rarely; synchronise so no in-flight kernel still uses the old buffer
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_doubles = 0;
    size_t n = std::max<size_t>(doubles, 4096);
    SBP_CUDA(cudaMalloc(&ctx->scratch, n * sizeof(double)));
    ctx->scratch_doubles = n;
    return SBP_OK;
}

int32_t launch_dense_small_ex(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                              double beta, double* energy, double* energy_sum, const double* cfrag,
                              const Sat1D* sat);
int32_t launch_sparse_ex(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                         double beta, const sbp_op* adv, const double* g, int64_t g_stride, const double* colscale,
                         const Sat1D* sat);
int32_t launch_adv1d_epilogue(sbp_ctx* ctx, const sbp_op* adv, double* dU, const double* U, int64_t B,
                              const double* bc, int64_t bc_stride, double* quantities, bool apply_sat);
static inline bool is_small(const sbp_op* op) {
    return op->variant == SBP_KERNEL_DMMA_SMEM || op->variant == SBP_KERNEL_DMMA_SMEM_SYM;
}
__host__ __device__ inline int phys_k_host(int ks, int t) { return 8 * (ks >> 1) + 2 * t + (ks & 1); }
// fragment layouts of dense_small.cu: node of (k-step, lane-in-group) and output row of (n-tile, column)
static inline int frag_k(int layout, int ks, int t) { return layout == 2 ? phys_k_host(ks, t) : 4 * ks + t; }
static inline int frag_row(int layout, int nt, int n) { return layout == 2 ? 8 * nt + n : 8 * nt + (n >> 1) + 4 * (n & 1); }

template <typename T>
static int32_t upload(T** dptr, const std::vector<T>& h) {
    *dptr = nullptr;
    if (h.empty()) return SBP_OK;
    SBP_CUDA(cudaMalloc(dptr, h.size() * sizeof(T)));
    SBP_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SBP_OK;
}

// Build device images of G dense N x N operators given column-major host matrices (leading dimension ldd).
static int32_t build_dense(sbp_op* op, int N, int G, const double* h_D, int ldd, const double* h_w, int flags) {
    sbp_ctx* ctx = op->ctx;
    op->N = N;
    op->G = G;
    op->nnz = (int64_t)G * N * N;
    const bool small = N <= kMaxSmallN && !(flags & SBP_DENSE_FORCE_GEMM);
    const size_t nn = (size_t)N * N;
    // row-major table (+ column-major copy for the generic table kernel when small)
    std::vector<double> rm(G * nn * (small ? 2 : 1));
    for (int g = 0; g < G; g++) {
        const double* Dg = h_D + (size_t)g * ldd * N;
        for (int i = 0; i < N; i++)
            for (int k = 0; k < N; k++) rm[g * nn + (size_t)i * N + k] = Dg[(size_t)k * ldd + i];
        if (small)
            for (int k = 0; k < N; k++)
                for (int i = 0; i < N; i++) rm[G * nn + g * nn + (size_t)k * N + i] = Dg[(size_t)k * ldd + i];
    }
    int32_t rc = upload(&op->d_D, rm);
    if (rc) return rc;
    if (h_w) {
        std::vector<double> w(h_w, h_w + (size_t)G * N);
        rc = upload(&op->d_w, w);
        if (rc) return rc;
    }
    if (small) {
        const int NT = (N + 7) / 8, KS = 2 * NT;
        op->NT = NT;
        const size_t FR = (size_t)KS * NT * 32;
        if ((int64_t)(G * FR * 8 + G * KS * 32 + 4096) <= ctx->smem_optin) {
            // fragment layout (dense_small.cu): 2 = paired k order / natural rows (N % 8 == 0, 128-bit accesses),
            //                                   1 = natural k order / interleaved rows (any N, 64-bit accesses)
            op->layout = (N % 8 == 0) ? 2 : 1;
            const int layout = op->layout;
            std::vector<double> frag(G * FR, 0.0), wfrag;
            size_t nonzero_blocks = 0;
            int band = 0;  // max |k-pair - n-tile| over the non-zero fragments
            for (int g = 0; g < G; g++) {
                const double* Dg = h_D + (size_t)g * ldd * N;
                for (int ks = 0; ks < KS; ks++)
                    for (int nt = 0; nt < NT; nt++) {
                        bool any = false;
                        for (int lane = 0; lane < 32; lane++) {
                            const int row = frag_row(layout, nt, lane >> 2), col = frag_k(layout, ks, lane & 3);
                            if (row < N && col < N) {
                                const double v = Dg[(size_t)col * ldd + row];
                                frag[g * FR + ((size_t)ks * NT + nt) * 32 + lane] = v;
                                any = any || (v != 0.0);
                            }
                        }
                        if (any) {
                            band = std::max(band, std::abs((ks >> 1) - nt));
                            nonzero_blocks++;
                        }
                    }
            }
            rc = upload(&op->d_frag, frag);
            if (rc) return rc;
            if (h_w) {
                wfrag.assign((size_t)G * KS * 4, 0.0);
                for (int g = 0; g < G; g++)
                    for (int ks = 0; ks < KS; ks++)
                        for (int t = 0; t < 4; t++) {
                            const int k = frag_k(layout, ks, t);
                            if (k < N) wfrag[(size_t)g * KS * 4 + ks * 4 + t] = h_w[(size_t)g * N + k];
                        }
                rc = upload(&op->d_wfrag, wfrag);
                if (rc) return rc;
            }
            // banded operators: fragments with |k-pair - n-tile| > band are structurally zero and are not multiplied
            const double zero_frac = 1.0 - (double)nonzero_blocks / ((double)G * KS * NT);
            op->zero_frac = zero_frac;
            const double kept = std::min(1.0, (2.0 * band + 1.0) / NT);  // fraction of fragments inside the band
            if (band >= 1 && band <= 2 && kept <= 0.75 && NT >= (band == 1 ? 4 : 6) && !(flags & SBP_DENSE_NO_SKIP))
                op->band = band;
            if (band == 0 && NT >= 4 && !(flags & SBP_DENSE_NO_SKIP)) op->band = 1;  // block diagonal: use band 1
            op->variant = SBP_KERNEL_DMMA_SMEM;
            // ---- "layout 4" image for the 256-bit kernel (dense_small_v4.cu): permuted k and output-row orders ----
            if (N % 16 == 0) {
                auto kap = [](int ks, int t) { return 16 * (ks >> 2) + 4 * t + (ks & 3); };
                auto rho = [](int nt, int n) { return 16 * (nt >> 1) + 4 * (n >> 1) + 2 * (nt & 1) + (n & 1); };
                std::vector<double> f4(G * FR, 0.0);
                for (int g = 0; g < G; g++) {
                    const double* Dg = h_D + (size_t)g * ldd * N;
                    for (int ks = 0; ks < KS; ks++)
                        for (int nt = 0; nt < NT; nt++)
                            for (int lane = 0; lane < 32; lane++)
                                f4[g * FR + ((size_t)ks * NT + nt) * 32 + lane] =
                                    Dg[(size_t)kap(ks, lane & 3) * ldd + rho(nt, lane >> 2)];
                }
                rc = upload(&op->d_frag4, f4);
                if (rc) return rc;
                if (h_w) {
                    std::vector<double> w4((size_t)G * KS * 4);
                    for (int g = 0; g < G; g++)
                        for (int ks = 0; ks < KS; ks++)
                            for (int t = 0; t < 4; t++) w4[(size_t)g * KS * 4 + ks * 4 + t] = h_w[(size_t)g * N + kap(ks, t)];
                    rc = upload(&op->d_wfrag4, w4);
                    if (rc) return rc;
                }
            }
            // ---- centro-antisymmetric operators: even/odd split, half the FP64 work (dense_small_sym.cu) ----
            if (G == 1 && N % 16 == 0 && !(flags & SBP_DENSE_NO_SYMMETRY)) {
                auto at = [&](int i, int k) { return h_D[(size_t)k * ldd + i]; };
                double dev = 0.0, amax = 0.0;
                for (int i = 0; i < N; i++)
                    for (int k = 0; k < N; k++) {
                        dev = std::max(dev, std::fabs(at(i, k) + at(N - 1 - i, N - 1 - k)));
                        amax = std::max(amax, std::fabs(at(i, k)));
                    }
                // centro-antisymmetric up to the rounding of its construction (a barycentric Legendre matrix as
                // PolynomialBases.jl builds it: deviation ~1e-16 |D|max): the antisymmetric part (D - J D J)/2 is used.
                // SBP_DENSE_SYM_BITWISE restricts the fast path to matrices that have the symmetry bitwise.
                const bool exact = dev == 0.0;
                const bool approx = !(flags & SBP_DENSE_SYM_BITWISE) && dev <= 256.0 * 2.220446049250313e-16 * amax;
                // the even/odd split halves the work; skipping zero fragments wins when more than half are zero
                const bool prefer_skip = op->band > 0 && (2.0 * op->band + 1.0) / NT < 0.5;
                if ((exact || approx) && !prefer_skip) {
                    op->band = 0;
                    const int h = N / 2, NTH = N / 16, KSH = 2 * NTH;
                    const size_t FH = (size_t)KSH * NTH * 32;
                    std::vector<double> fs(2 * FH, 0.0);
                    // antisymmetric part (== D when the symmetry is exact)
                    auto da = [&](int i, int k) { return exact ? at(i, k) : 0.5 * (at(i, k) - at(N - 1 - i, N - 1 - k)); };
                    for (int ks = 0; ks < KSH; ks++)
                        for (int nt = 0; nt < NTH; nt++)
                            for (int lane = 0; lane < 32; lane++) {
                                const int i = nt * 8 + (lane >> 2), j = phys_k_host(ks, lane & 3);
                                (void)h;
                                const size_t o = ((size_t)ks * NTH + nt) * 32 + lane;
                                fs[o] = 0.5 * (da(i, j) + da(i, N - 1 - j));       // E
                                fs[FH + o] = 0.5 * (da(i, j) - da(i, N - 1 - j));  // O
                            }
                    rc = upload(&op->d_frag_sym, fs);
                    if (rc) return rc;
                    if (h_w) {
                        std::vector<double> ws((size_t)2 * KSH * 4);
                        for (int ks = 0; ks < KSH; ks++)
                            for (int t = 0; t < 4; t++) {
                                const int j = phys_k_host(ks, t);
                                ws[ks * 4 + t] = h_w[j];
                                ws[(size_t)KSH * 4 + ks * 4 + t] = h_w[N - 1 - j];
                            }
                        rc = upload(&op->d_wsym, ws);
                        if (rc) return rc;
                    }
                    op->sym = true;
                    op->variant = SBP_KERNEL_DMMA_SMEM_SYM;
                }
            }
        } else {
            op->variant = SBP_KERNEL_TABLE;  // table too large for shared memory: generic kernel only
        }
    } else {
        op->variant = SBP_KERNEL_DGEMM_TILED;
    }
    return SBP_OK;
}

// Row-blocked SELL-32 (see sparse.cu): a lane owns R consecutive rows and walks the union of their column sets.
// R is chosen to minimise the estimated LSU wavefronts (16 per gathered column + 1 + 2R for the operator entry).
static inline int host_node_slot(int c, int logR) {
    const int R = 1 << logR;
    const int cb = c >> logR, r = c & (R - 1);
    return ((cb >> 1) << (logR + 1)) | (r << 1) | (cb & 1);
}

static int32_t build_sell(sbp_op* op, int N, const std::vector<int>& rowptr, const std::vector<int>& col,
                          const std::vector<double>& val) {
    int best_logR = 0;
    double best_cost = 1e300;
    std::vector<std::vector<int>> best_union;
    const int forced = op->ctx->tune[T_SPARSE_LOGR];  // tuning/debug knob: force rows-per-lane = 1 << value
    for (int logR = 0; logR <= 2; logR++) {
        if (forced >= 0 && forced != logR) continue;
        const int R = 1 << logR;
        const int nrb = (N + R - 1) / R;
        std::vector<std::vector<int>> uni(nrb);
        for (int rb = 0; rb < nrb; rb++) {
            std::vector<int>& u = uni[rb];
            for (int r = rb * R; r < std::min(N, rb * R + R); r++)
                for (int j = rowptr[r]; j < rowptr[r + 1]; j++) u.push_back(col[j]);
            std::sort(u.begin(), u.end());
            u.erase(std::unique(u.begin(), u.end()), u.end());
        }
        double cost = 0.0;
        for (int s = 0; s * 32 < nrb; s++) {
            size_t w = 0;
            for (int rb = s * 32; rb < std::min(nrb, s * 32 + 32); rb++) w = std::max(w, uni[rb].size());
            cost += (double)w * (17.0 + 2.0 * R);
        }
        if (R == 4) cost *= 1.2;  // measured: 32 accumulators + operator prefetch cost occupancy (158 vs 124 registers)
        // a tile is worked on by one warp per slice: too few slices leave the CTA's other warps idle (N = 101, R = 4
        // is a single slice and ran 30% slower than R = 1 with its 4 slices)
        const int nsl = (nrb + 31) / 32;
        if (nsl < 4) cost *= 4.0 / nsl;
        if (cost < best_cost * 0.97) {  // prefer the smaller R unless the gain is real
            best_cost = cost;
            best_logR = logR;
            best_union.swap(uni);
        }
    }
    const int logR = best_logR, R = 1 << logR;
    const int nrb = (N + R - 1) / R;
    const int n_slices = (nrb + 31) / 32;
    std::vector<int> slice_ptr(n_slices + 1, 0);
    int maxw = 0;
    for (int s = 0; s < n_slices; s++) {
        int w = 0;
        for (int rb = s * 32; rb < std::min(nrb, s * 32 + 32); rb++) w = std::max(w, (int)best_union[rb].size());
        w += w & 1;  // even width: the kernel consumes two union columns per step
        slice_ptr[s + 1] = slice_ptr[s] + w;
        maxw = std::max(maxw, w);
    }
    const size_t entries = (size_t)slice_ptr[n_slices];
    std::vector<int> c_slot(entries * 32), c_node(entries * 32);
    std::vector<double> v(entries * R * 32, 0.0);
    for (int s = 0; s < n_slices; s++) {
        const int w = slice_ptr[s + 1] - slice_ptr[s];
        for (int l = 0; l < 32; l++) {
            const int rb = s * 32 + l;
            const int pad_node = std::min(rb * R, N - 1);  // padding: value 0 times a valid column
            const std::vector<int>* u = rb < nrb ? &best_union[rb] : nullptr;
            for (int j = 0; j < w; j++) {
                const size_t e = (size_t)slice_ptr[s] + j;
                const int node = (u && j < (int)u->size()) ? (*u)[j] : pad_node;
                c_node[e * 32 + l] = node;
                c_slot[e * 32 + l] = host_node_slot(node, logR);
            }
            if (!u) continue;
            for (int r = 0; r < R; r++) {
                const int row = rb * R + r;
                if (row >= N) break;
                for (int jj = rowptr[row]; jj < rowptr[row + 1]; jj++) {
                    const int j = (int)(std::lower_bound(u->begin(), u->end(), col[jj]) - u->begin());
                    v[(((size_t)slice_ptr[s] + j) * R + r) * 32 + l] += val[jj];  // duplicates accumulate
                }
            }
        }
    }
    op->logR = logR;
    op->n_slices = n_slices;
    op->max_row_nnz = maxw;
    int32_t rc = upload(&op->d_slice_ptr, slice_ptr);
    if (rc) return rc;
    rc = upload(&op->d_sell_col, c_slot);
    if (rc) return rc;
    rc = upload(&op->d_col, c_node);
    if (rc) return rc;
    return upload(&op->d_sell_val, v);
}

// Block-sparse DMMA image (bsr.cu).
//  * ROW BLOCKS of 8 rows = two QUADS of 4 consecutive rows (aligned to 4).  The second quad is chosen greedily among
//    the quads that share columns with the first one so that the union of the block's columns is smallest: for a
//    stencil-like operator on a 2-D cloud a 2 x 4 patch of nodes needs ~25% fewer fragments than 8 rows in a line; for
//    banded 1-D operators the partner is simply the next quad.  A lane of the C fragment owns an aligned row pair, so
//    the output is still written with 16-byte stores.
//  * FRAGMENTS: the union of a block's columns, 4 columns each (see the bank-conflict note below).
//  * CHUNKS of 4 row blocks define the window of TMA boxes (kBsrBX nodes) [blo_c, bhi_c) that must be resident while
//    the chunk is multiplied; nbox[d] is the ring size that lets the boxes of the next d chunks stream in meanwhile
//    (the stream continues across the tiles of a CTA, nbt boxes per tile).
static int32_t build_bsr(sbp_op* op, int N1, const std::vector<int>& rowptr1, const std::vector<int>& col1,
                         const std::vector<double>& val1) {
    // Odd N: the rows of the batch are not multiples of 16 bytes, which the TMA needs (global strides AND box origins:
    // an odd innermost coordinate is an illegal instruction, tools/tma_coord_test.cu).  The batch is then read as
    // [B/2][2N] (two instances per tensor row) and multiplied by the block-diagonal operator diag(D, D) of size 2N: the
    // even-N data path runs unchanged, row blocks that straddle the two halves are ordinary sparse row blocks.
    const bool paired = (N1 & 1) != 0;
    const int N = paired ? 2 * N1 : N1;
    std::vector<int> rowptr2, col2;
    std::vector<double> val2;
    if (paired) {
        rowptr2.assign(N + 1, 0);
        col2.reserve(2 * col1.size());
        val2.reserve(2 * val1.size());
        for (int h = 0; h < 2; h++)
            for (int r = 0; r < N1; r++) {
                for (int j = rowptr1[r]; j < rowptr1[r + 1]; j++) {
                    col2.push_back(col1[j] + h * N1);
                    val2.push_back(val1[j]);
                }
                rowptr2[h * N1 + r + 1] = (int)col2.size();
            }
    }
    const std::vector<int>& rowptr = paired ? rowptr2 : rowptr1;
    const std::vector<int>& col = paired ? col2 : col1;
    const std::vector<double>& val = paired ? val2 : val1;
    op->bsr_N = N;
    constexpr int CH = kBsrCH;
    constexpr int BX = kBsrBX;
    const int nq = (N + 3) / 4;
    auto union_of = [&](int q, std::vector<int>& u) {
        for (int r = 4 * q; r < std::min(N, 4 * q + 4); r++)
            for (int j = rowptr[r]; j < rowptr[r + 1]; j++) u.push_back(col[j]);
    };
    auto uniq = [](std::vector<int>& u) {
        std::sort(u.begin(), u.end());
        u.erase(std::unique(u.begin(), u.end()), u.end());
    };
    std::vector<std::vector<int>> qcols(nq);
    for (int q = 0; q < nq; q++) {
        union_of(q, qcols[q]);
        uniq(qcols[q]);
    }
    // greedy pairing of quads
    std::vector<char> used(nq, 0);
    std::vector<int> rows;  // 2 per block: first row of quad A, first row of quad B (or >= N: none)
    const int none = 4 * nq + 4;
    const bool pair_far = op->ctx->tune[T_BSR_NO_PAIRING] == 0;
    for (int qa = 0; qa < nq; qa++) {
        if (used[qa]) continue;
        used[qa] = 1;
        int best = -1;
        size_t best_size = ~size_t(0);
        std::vector<int> cand;
        if (pair_far)
            for (int cidx : qcols[qa]) cand.push_back(cidx / 4);
        if (qa + 1 < nq) cand.push_back(qa + 1);
        uniq(cand);
        std::vector<int> merged;
        for (int qb : cand) {
            if (used[qb]) continue;
            merged.clear();
            std::set_union(qcols[qa].begin(), qcols[qa].end(), qcols[qb].begin(), qcols[qb].end(),
                           std::back_inserter(merged));
            const size_t sz = (merged.size() + 3) / 4;  // fragments
            if (sz < best_size || (sz == best_size && std::abs(qb - qa) < std::abs(best - qa))) {
                best_size = sz;
                best = qb;
            }
        }
        if (best < 0)  // no unassigned neighbour: take the next free quad so that blocks stay full
            for (int qb = qa + 1; qb < nq; qb++)
                if (!used[qb]) {
                    best = qb;
                    break;
                }
        rows.push_back(4 * qa);
        if (best >= 0) {
            used[best] = 1;
            rows.push_back(4 * best);
        } else {
            rows.push_back(none);
        }
    }
    const int nrb = (int)rows.size() / 2, nch = (nrb + CH - 1) / CH;
    auto block_row = [&](int rb, int g) { return (g < 4 ? rows[2 * rb] : rows[2 * rb + 1] - 4) + g; };
    constexpr int PS = 4;  // ring codes per fragment
    std::vector<int> rbf(nrb + 1, 0), pos, diag(8 * (size_t)nrb, -1);
    std::vector<double> frag;
    std::vector<int> cmin(nch, N), cmax(nch, 0);
    for (int rb = 0; rb < nrb; rb++) {
        std::vector<int> u;
        const int c = rb / CH;
        for (int g = 0; g < 8; g++) {
            const int r = block_row(rb, g);
            if (r >= N) continue;
            for (int j = rowptr[r]; j < rowptr[r + 1]; j++) u.push_back(col[j]);
            u.push_back(r);  // the diagonal entry always has a place in the image (SAT and stage-update terms are folded into it)
            // the block's own rows are part of the chunk's window (the SAT epilogues read u_j from the ring)
            cmin[c] = std::min(cmin[c], r);
            cmax[c] = std::max(cmax[c], r + 1);
        }
        uniq(u);
        // Fragments of 4 columns.  The A-operand gather of a fragment is bank-conflict free iff its 4 columns differ
        // mod 4 (rows of a box are 36 = 4 (mod 16) doubles apart), so columns are dealt from 4 residue buckets, the
        // fullest first; a fragment short of columns repeats one of its own (same address => broadcast, value 0).
        const int nf = ((int)u.size() + 3) / 4;
        const size_t f0 = frag.size() / 32;
        frag.resize((f0 + nf) * 32, 0.0);
        pos.resize((f0 + nf) * PS, 0);
        std::vector<int> where(u.size(), 0);  // union index -> 4 * fragment + t
        {
            std::vector<int> bucket[4];
            for (int k = (int)u.size() - 1; k >= 0; k--) bucket[u[k] & 3].push_back(k);  // pop_back yields ascending
            for (int f = 0; f < nf; f++) {
                int taken = 0, cols[4];
                int order[4] = {0, 1, 2, 3};
                std::sort(order, order + 4, [&](int x, int y) { return bucket[x].size() > bucket[y].size(); });
                for (int pass = 0; pass < 4 && taken < 4; pass++)  // later passes: fewer than 4 residue classes left
                    for (int q = 0; q < 4 && taken < 4; q++) {
                        std::vector<int>& bk = bucket[order[q]];
                        if (!bk.empty()) {
                            cols[taken++] = bk.back();
                            bk.pop_back();
                        }
                    }
                for (int t = 0; t < 4; t++) {
                    const int k = cols[t < taken ? t : 0];
                    pos[(f0 + f) * PS + t] = u[k];
                    if (t < taken) where[k] = 4 * f + t;
                }
            }
        }
        for (int g = 0; g < 8; g++) {
            const int r = block_row(rb, g);
            if (r >= N) continue;
            for (int j = rowptr[r]; j < rowptr[r + 1]; j++) {
                const int k = (int)(std::lower_bound(u.begin(), u.end(), col[j]) - u.begin());
                const int w = where[k];
                frag[(f0 + w / 4) * 32 + (size_t)g * 4 + (w & 3)] += val[j];  // lane = 4*g + t; duplicates accumulate
            }
            const int w = where[(int)(std::lower_bound(u.begin(), u.end(), r) - u.begin())];
            diag[8 * (size_t)rb + g] = (int)((f0 + w / 4) * 32 + (size_t)g * 4 + (w & 3));
        }
        rbf[rb + 1] = (int)(f0 + nf);
        if (!u.empty()) {
            cmin[c] = std::min(cmin[c], u.front());
            cmax[c] = std::max(cmax[c], u.back() + 1);
        }
    }
    // windows in units of TMA boxes: boxes [0, bhi[c]) resident for chunk c, none below blo[c] needed later
    std::vector<int> blo(nch), bhi(nch);
    for (int c = 0, m = 0; c < nch; c++) {
        m = std::max(m, cmax[c]);
        bhi[c] = (m + BX - 1) / BX;
    }
    for (int c = nch - 1, m = N; c >= 0; c--) {
        m = std::min(m, cmin[c]);
        blo[c] = m / BX;
    }
    // pos3: the same codes as byte offsets into the ring for every tile size MT = 1..8 (box = 8 MT instances x BX nodes)
    std::vector<int> pos3(8 * pos.size(), 0);
    int maxnf = 0;
    for (int rb = 0; rb < nrb; rb++) {
        maxnf = std::max(maxnf, rbf[rb + 1] - rbf[rb]);
        for (int f = rbf[rb] * PS; f < rbf[rb + 1] * PS; f++) {
            const int n = pos[f];
            const int brel = n / BX - blo[rb / CH];
            pos[f] = (brel << 8) | (n % BX);
            for (int mt = 1; mt <= 8; mt++) pos3[(size_t)(mt - 1) * pos.size() + f] = brel * (8 * mt * BX * 8) + (n % BX) * 8;
        }
    }
    op->bsr_maxnf = maxnf;
    // ring sizes: in the virtual box stream (tile k occupies boxes [k*nbt, (k+1)*nbt)) the boxes of item s may only
    // overwrite slots that no item >= s - d still needs
    const int nbt = (N + BX - 1) / BX;
    int maxf = 1;
    std::vector<int> rbmeta(4 * (size_t)nrb), chmeta(4 * (size_t)nch);
    for (int c = 0; c < nch; c++) {
        const int fa = rbf[c * CH], fb = rbf[std::min(nrb, c * CH + CH)];
        maxf = std::max(maxf, fb - fa);
        chmeta[4 * c + 0] = fa, chmeta[4 * c + 1] = fb, chmeta[4 * c + 2] = blo[c], chmeta[4 * c + 3] = bhi[c];
    }
    for (int rb = 0; rb < nrb; rb++) {
        rbmeta[4 * rb + 0] = rbf[rb] - rbf[(rb / CH) * CH];
        rbmeta[4 * rb + 1] = rbf[rb + 1] - rbf[rb];
        rbmeta[4 * rb + 2] = rows[2 * rb];
        rbmeta[4 * rb + 3] = rows[2 * rb + 1];
    }
    for (int d = 0; d < kBsrDepths; d++) {
        long long need = 1;
        const int items = nch * (d + 2);
        for (int s = 0; s < items; s++) {
            const int sb = std::max(0, s - d);
            const long long vhi = (long long)(s / nch) * nbt + bhi[s % nch];
            const long long vlo = (long long)(sb / nch) * nbt + blo[sb % nch];
            need = std::max(need, vhi - vlo);
        }
        op->bsr_nbox[d] = (int)need;
    }
    // per-row-block records read by the consumer warps, one array per pipeline depth d (the ring size differs):
    // {first fragment relative to the chunk, fragments, first row of quad A, of quad B | ring slot of the chunk's lowest
    // box (blo % nbox[d]), blo, 0, 0}; padded to whole chunks with inactive records (rows >= N, no fragments)
    std::vector<int> item((size_t)kBsrDepths * nch * CH * 8, 0);
    int pairs = 0;
    for (int d = 0; d < kBsrDepths; d++)
        for (int rb = 0; rb < nch * CH; rb++) {
            int* e = &item[((size_t)d * nch * CH + rb) * 8];
            if (rb < nrb) {
                for (int k = 0; k < 4; k++) e[k] = rbmeta[4 * rb + k];
                e[4] = blo[rb / CH] % op->bsr_nbox[d];
                e[5] = blo[rb / CH];
                // pair flag (bsr3.cu): blocks 2j, 2j + 1 of a chunk whose quads are x-neighbours (A next to A, B next to B) are staged
                // side by side and leave as two {8 x instances} boxes instead of four {4 x instances}: 1 = left block, 2 = right block
                {
                    const int l = rb & ~1, r = l + 1;
                    const bool self_l = rows[2 * l + 1] == rows[2 * l] + 4;
                    if (r < nrb && !self_l && rows[2 * r + 1] != rows[2 * r] + 4 && rows[2 * r] == rows[2 * l] + 4 &&
                        rows[2 * r + 1] == rows[2 * l + 1] + 4 && rows[2 * r + 1] + 4 <= N && rows[2 * l + 1] < N)
                        e[7] = rb == l ? 1 : 2, pairs += d == 0;
                }
                // rows of the block that carry a 1D SAT: bit k = block row k is node 0 of its instance, bit 8 + k = node N1 - 1
                for (int k = 0; k < 8; k++) {
                    const int r = block_row(rb, k);
                    if (r >= N) continue;
                    const int n = r % N1;
                    if (n == 0) e[6] |= 1 << k;
                    if (n == N1 - 1) e[6] |= 1 << (8 + k);
                }
            } else {
                e[2] = 1 << 30, e[3] = (1 << 30) + 4;
            }
        }
    if (const char* e = getenv("SBP_BSR_STATS"); e && *e == '1') {  // debug aid (operator creation, not a launch path)
        long long slot[CH] = {}, sched[4] = {}, sum_max = 0;
        for (int c = 0; c < nch; c++) {
            long long sc[4] = {};
            for (int k = 0; k < CH; k++)
                if (c * CH + k < nrb) slot[k] += rbmeta[4 * (c * CH + k) + 1], sc[k & 3] += rbmeta[4 * (c * CH + k) + 1];
            for (int q = 0; q < 4; q++) sched[q] += sc[q];
            sum_max += *std::max_element(sc, sc + 4);
        }
        fprintf(stderr, "bsr stats: N %d, %d row blocks, %d chunks, %lld fragments (%.2f per block), max per block %d\n  per slot:", N, nrb,
                nch, (long long)frag.size() / 32, (double)(frag.size() / 32) / nrb, maxnf);
        for (int k = 0; k < CH; k++) fprintf(stderr, " %lld", slot[k]);
        fprintf(stderr, "\n  per scheduler: %lld %lld %lld %lld, sum over chunks of the busiest scheduler %lld (mean %.1f)\n", sched[0], sched[1],
                sched[2], sched[3], sum_max, (double)(frag.size() / 32) / 4);
    }
    op->bsr_maxf = maxf;
    op->bsr_nrb = nrb;
    op->bsr_rows = rows;
    op->bsr_diag = diag;
    op->bsr_pairs = pairs;
    op->bsr_frags = (int64_t)frag.size() / 32;
    if (frag.empty()) {
        frag.assign(32, 0.0);
        pos.assign(PS, 0);
    }
    int32_t rc = upload(&op->d_bsr_val, frag);

    if (!rc) rc = upload(&op->d_bsr_pos, pos);
    if (pos3.empty()) pos3.assign(8 * PS, 0);
    if (!rc) rc = upload(&op->d_bsr_pos3, pos3);
    if (!rc) rc = upload(&op->d_bsr_rbmeta, rbmeta);
    if (!rc) rc = upload(&op->d_bsr_chmeta, chmeta);
    if (!rc) rc = upload(&op->d_bsr_item, item);
    return rc;
}

// ---- NCCL through dlopen (no link-time dependency; torch's bundled libnccl is reused when loaded) -----
struct Id128 {
    char b[128];
};
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int32_t load_nccl() {
    if (g_nccl.handle) return SBP_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    SBP_REQUIRE(h != nullptr, SBP_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    SBP_REQUIRE(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy, SBP_ENCCL,
                "libnccl is missing required symbols");
    g_nccl.handle = h;
    return SBP_OK;
}

// ---- host-buffer pipeline -----------------------------------------------------------------------------
// Chunks of instances travel H2D -> kernel -> D2H over three streams with two staging slots.  `run` launches the hot-path
// call for nb instances starting at instance `done` on device buffers (dIn -> dOut) on the ctx stream.
template <typename Run>
static int32_t host_pipeline(sbp_ctx* ctx, int N, int G, double* h_out, const double* h_in, int64_t B, bool read_out,
                             int64_t chunk_instances, Run run) {
    if (chunk_instances <= 0) chunk_instances = std::max<int64_t>(1, ((int64_t)32 << 20) / ((int64_t)N * 8));
    // keep chunks a multiple of 128*G so that table periodicity and kernel tiling are preserved
    const int64_t q = 128 * (int64_t)G;
    chunk_instances = std::max<int64_t>(q, (chunk_instances / q) * q);
    chunk_instances = std::min<int64_t>(chunk_instances, ((B + q - 1) / q) * q);
    const size_t cbytes = (size_t)chunk_instances * N * sizeof(double);
    if (ctx->pipe_bytes < cbytes) {
        SBP_CUDA(cudaDeviceSynchronize());
        for (int s = 0; s < 2; s++)
            for (int k = 0; k < 2; k++) {
                if (ctx->pipe_dev[s][k]) cudaFree(ctx->pipe_dev[s][k]);
                ctx->pipe_dev[s][k] = nullptr;
            }
        ctx->pipe_bytes = 0;
        for (int s = 0; s < 2; s++)
            for (int k = 0; k < 2; k++) SBP_CUDA(cudaMalloc(&ctx->pipe_dev[s][k], cbytes));
        ctx->pipe_bytes = cbytes;
    }
    cudaStream_t compute = ctx->stream;
    // the pipeline must start after work already queued on the compute stream
    SBP_CUDA(cudaEventRecord(ctx->pipe_ev[0][2], compute));
    SBP_CUDA(cudaStreamWaitEvent(ctx->s_in, ctx->pipe_ev[0][2], 0));
    SBP_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->pipe_ev[0][2], 0));
    int64_t done = 0;
    int c = 0;
    enum { EV_IN = 0, EV_COMP = 1, EV_OUT = 2 };
    bool used[2] = {false, false};
    while (done < B) {
        const int s = c & 1;
        const int64_t nb = std::min<int64_t>(chunk_instances, B - done);
        const size_t bytes = (size_t)nb * N * sizeof(double);
        double* dIn = static_cast<double*>(ctx->pipe_dev[s][0]);
        double* dOut = static_cast<double*>(ctx->pipe_dev[s][1]);
        if (used[s]) {
            SBP_CUDA(cudaStreamWaitEvent(ctx->s_in, ctx->pipe_ev[s][EV_COMP], 0));  // dIn[s] consumed
            if (read_out) SBP_CUDA(cudaStreamWaitEvent(ctx->s_in, ctx->pipe_ev[s][EV_OUT], 0));
        }
        SBP_CUDA(cudaMemcpyAsync(dIn, h_in + done * N, bytes, cudaMemcpyHostToDevice, ctx->s_in));
        if (read_out) SBP_CUDA(cudaMemcpyAsync(dOut, h_out + done * N, bytes, cudaMemcpyHostToDevice, ctx->s_in));
        SBP_CUDA(cudaEventRecord(ctx->pipe_ev[s][EV_IN], ctx->s_in));
        SBP_CUDA(cudaStreamWaitEvent(compute, ctx->pipe_ev[s][EV_IN], 0));
        if (used[s]) SBP_CUDA(cudaStreamWaitEvent(compute, ctx->pipe_ev[s][EV_OUT], 0));  // dOut[s] drained
        int32_t rc = run(dOut, dIn, nb, done);
        if (rc) return rc;
        SBP_CUDA(cudaEventRecord(ctx->pipe_ev[s][EV_COMP], compute));
        SBP_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->pipe_ev[s][EV_COMP], 0));
        SBP_CUDA(cudaMemcpyAsync(h_out + done * N, dOut, bytes, cudaMemcpyDeviceToHost, ctx->s_out));
        SBP_CUDA(cudaEventRecord(ctx->pipe_ev[s][EV_OUT], ctx->s_out));
        used[s] = true;
        done += nb;
        c++;
    }
    SBP_CUDA(cudaStreamSynchronize(ctx->s_out));
    SBP_CUDA(cudaStreamSynchronize(compute));
    return SBP_OK;
}

// grow-only device buffer owned by the ctx (boundary data / work vectors of the *_host and solve entry points)
int32_t ensure_aux(sbp_ctx* ctx, int slot, size_t bytes, void** out) {
    if (ctx->aux_bytes[slot] < bytes) {
        SBP_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->aux[slot]) cudaFree(ctx->aux[slot]);
        ctx->aux[slot] = nullptr, ctx->aux_bytes[slot] = 0;
        SBP_CUDA(cudaMalloc(&ctx->aux[slot], bytes));
        ctx->aux_bytes[slot] = bytes;
    }
    *out = ctx->aux[slot];
    return SBP_OK;
}

}  // namespace sbp

using namespace sbp;

extern "C" {

int32_t sbp_version(void) { return 100; }  // 0.1.0
const char* sbp_last_error(void) { return g_last_error.c_str(); }

int32_t sbp_device_count(int32_t* n) {
    SBP_REQUIRE(n, SBP_EINVAL, "null output");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        cudaGetLastError();
        c = 0;
    }
    *n = c;
    return SBP_OK;
}

int32_t sbp_ctx_create(int32_t device, sbp_ctx** out) {
    SBP_REQUIRE(out, SBP_EINVAL, "null output");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libsbp_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return SBP_ECUDA;
    }
    SBP_REQUIRE(device >= 0 && device < count, SBP_EINVAL, "device %d out of range [0,%d)", device, count);
    SBP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SBP_CUDA(cudaGetDeviceProperties(&prop, device));
    SBP_REQUIRE(prop.major >= 10, SBP_EUNSUPPORTED,
                "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    sbp_ctx* ctx = new sbp_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_bytes = prop.l2CacheSize;
    ctx->smem_optin = (int64_t)prop.sharedMemPerBlockOptin;
    ctx->global_bytes = (int64_t)prop.totalGlobalMem;
    SBP_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    SBP_CUDA(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    SBP_CUDA(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    tune_from_env(ctx);
    for (int s = 0; s < 2; s++)
        for (int k = 0; k < 3; k++) SBP_CUDA(cudaEventCreateWithFlags(&ctx->pipe_ev[s][k], cudaEventDisableTiming));
    *out = ctx;
    return SBP_OK;
}

int32_t sbp_ctx_set_stream(sbp_ctx* ctx, void* cuda_stream) {
    SBP_REQUIRE(ctx, SBP_EINVAL, "null ctx");
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return SBP_OK;
}
int32_t sbp_ctx_get_stream(sbp_ctx* ctx, void** cuda_stream) {
    SBP_REQUIRE(ctx && cuda_stream, SBP_EINVAL, "null argument");
    *cuda_stream = ctx->stream;
    return SBP_OK;
}
int32_t sbp_ctx_sync(sbp_ctx* ctx) {
    SBP_REQUIRE(ctx, SBP_EINVAL, "null ctx");
    SBP_CUDA(cudaSetDevice(ctx->device));
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    return SBP_OK;
}
int32_t sbp_ctx_destroy(sbp_ctx* ctx) {
    if (!ctx) return SBP_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->stage_buf) cudaFree(ctx->stage_buf);
    for (void* a : ctx->aux)
        if (a) cudaFree(a);
    for (int s = 0; s < 2; s++) {
        for (int k = 0; k < 2; k++)
            if (ctx->pipe_dev[s][k]) cudaFree(ctx->pipe_dev[s][k]);
        for (int k = 0; k < 3; k++)
            if (ctx->pipe_ev[s][k]) cudaEventDestroy(ctx->pipe_ev[s][k]);
    }
    cudaStreamDestroy(ctx->own_stream);
    cudaStreamDestroy(ctx->s_in);
    cudaStreamDestroy(ctx->s_out);
    delete ctx;
    return SBP_OK;
}
int32_t sbp_ctx_device_info(sbp_ctx* ctx, int32_t* sm_count, int64_t* l2_bytes, int64_t* smem_optin,
                            int64_t* global_bytes) {
    SBP_REQUIRE(ctx, SBP_EINVAL, "null ctx");
    if (sm_count) *sm_count = ctx->sm_count;
    if (l2_bytes) *l2_bytes = ctx->l2_bytes;
    if (smem_optin) *smem_optin = ctx->smem_optin;
    if (global_bytes) *global_bytes = ctx->global_bytes;
    return SBP_OK;
}
int64_t sbp_ctx_launch_count(sbp_ctx* ctx) { return ctx ? ctx->launches : -1; }

static int tune_index(const char* name) {
    for (int k = 0; k < T_COUNT; k++)
        if (strcmp(name, kTuneSpec[k].name) == 0 || strcmp(name, kTuneSpec[k].env) == 0) return k;
    return -1;
}
int32_t sbp_ctx_set_option(sbp_ctx* ctx, const char* name, int32_t value) {
    SBP_REQUIRE(ctx && name, SBP_EINVAL, "null argument");
    const int k = tune_index(name);
    SBP_REQUIRE(k >= 0, SBP_EINVAL, "unknown option '%s'", name);
    ctx->tune[k] = value;
    return SBP_OK;
}
int32_t sbp_ctx_get_option(sbp_ctx* ctx, const char* name, int32_t* value) {
    SBP_REQUIRE(ctx && name && value, SBP_EINVAL, "null argument");
    const int k = tune_index(name);
    SBP_REQUIRE(k >= 0, SBP_EINVAL, "unknown option '%s'", name);
    *value = ctx->tune[k];
    return SBP_OK;
}

// ---- memory -------------------------------------------------------------------------------------
int32_t sbp_malloc(sbp_ctx* ctx, size_t bytes, void** dptr) {
    SBP_REQUIRE(ctx && dptr, SBP_EINVAL, "null argument");
    SBP_CUDA(cudaSetDevice(ctx->device));
    SBP_CUDA(cudaMalloc(dptr, bytes ? bytes : 16));
    return SBP_OK;
}
int32_t sbp_free(sbp_ctx* ctx, void* dptr) {
    SBP_REQUIRE(ctx, SBP_EINVAL, "null ctx");
    if (!dptr) return SBP_OK;
    SBP_CUDA(cudaSetDevice(ctx->device));
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    SBP_CUDA(cudaFree(dptr));
    return SBP_OK;
}
int32_t sbp_ctx_pci_bus_id(sbp_ctx* ctx, char* buf, int32_t len) {
    SBP_REQUIRE(ctx && buf && len >= 16, SBP_EINVAL, "sbp_ctx_pci_bus_id: need a ctx and a buffer of at least 16 bytes");
    SBP_CUDA(cudaDeviceGetPCIBusId(buf, len, ctx->device));
    for (char* c = buf; *c; c++)
        if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    return SBP_OK;
}

int32_t sbp_host_alloc(size_t bytes, void** hptr) {
    SBP_REQUIRE(hptr, SBP_EINVAL, "null output");
    SBP_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 16, cudaHostAllocDefault));
    return SBP_OK;
}
int32_t sbp_host_free(void* hptr) {
    if (hptr) SBP_CUDA(cudaFreeHost(hptr));
    return SBP_OK;
}
int32_t sbp_memcpy_h2d(sbp_ctx* ctx, void* dptr, const void* h_src, size_t bytes) {
    SBP_REQUIRE(ctx && (bytes == 0 || (dptr && h_src)), SBP_EINVAL, "null argument");
    SBP_CUDA(cudaMemcpyAsync(dptr, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SBP_OK;
}
int32_t sbp_memcpy_d2h(sbp_ctx* ctx, void* h_dst, const void* dptr, size_t bytes) {
    SBP_REQUIRE(ctx && (bytes == 0 || (dptr && h_dst)), SBP_EINVAL, "null argument");
    SBP_CUDA(cudaMemcpyAsync(h_dst, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    return SBP_OK;
}
int32_t sbp_memset(sbp_ctx* ctx, void* dptr, int32_t byte, size_t bytes) {
    SBP_REQUIRE(ctx && dptr, SBP_EINVAL, "null argument");
    SBP_CUDA(cudaMemsetAsync(dptr, byte, bytes, ctx->stream));
    return SBP_OK;
}

// ---- operators ----------------------------------------------------------------------------------
static sbp_op* new_op(sbp_ctx* ctx, int kind) {
    sbp_op* op = new sbp_op();
    op->ctx = ctx;
    op->kind = kind;
    return op;
}

int32_t sbp_op_destroy(sbp_op* op) {
    if (!op) return SBP_OK;
    cudaSetDevice(op->ctx->device);
    cudaStreamSynchronize(op->ctx->stream);
    void* ptrs[] = {op->d_D,         op->d_frag,     op->d_wfrag,    op->d_w,        op->d_frag_sym, op->d_rowptr,
                    op->d_col,       op->d_val,      op->d_slice_ptr, op->d_sell_col, op->d_sell_val, op->d_b,
                    op->d_mode,      op->d_bnd_node, op->d_bnd_tau,  op->d_a,        op->d_q_consts, op->d_wsym,
                    op->d_frag4,     op->d_wfrag4,   op->d_bsr_val,  op->d_bsr_pos,  op->d_bsr_rbmeta, op->d_bsr_chmeta,
                    op->d_rb_bdiag,  op->d_bsr_item, op->d_bsr_pos3, op->d_csr_rowptr, op->d_csr_col, op->d_csr_val,
                    op->d_bsr_val_sat, op->d_bsr_val_stage, op->d_bsr_diagmask};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete op;
    return SBP_OK;
}

int32_t sbp_op_dense_create(sbp_ctx* ctx, int32_t N, const double* h_D, int32_t ldd, const double* h_w,
                            double scale, int32_t flags, sbp_op** out) {
    SBP_REQUIRE(ctx && h_D && out, SBP_EINVAL, "null argument");
    SBP_REQUIRE(N >= 1 && ldd >= N, SBP_EDIM, "need N >= 1 and ldd >= N (N=%d, ldd=%d)", N, ldd);
    SBP_CUDA(cudaSetDevice(ctx->device));
    sbp_op* op = new_op(ctx, SBP_OP_DENSE);
    op->scale = scale;
    int32_t rc = build_dense(op, N, 1, h_D, ldd, h_w, flags);
    if (rc) {
        sbp_op_destroy(op);
        return rc;
    }
    *out = op;
    return SBP_OK;
}

int32_t sbp_op_table_create(sbp_ctx* ctx, int32_t N, int32_t G, const double* h_D, const double* h_w,
                            sbp_op** out) {
    SBP_REQUIRE(ctx && h_D && out, SBP_EINVAL, "null argument");
    SBP_REQUIRE(N >= 1 && G >= 1, SBP_EDIM, "need N >= 1 and G >= 1");
    SBP_REQUIRE(N <= kMaxSmallN, SBP_EUNSUPPORTED, "operator tables support N <= %d (got %d)", kMaxSmallN, N);
    SBP_CUDA(cudaSetDevice(ctx->device));
    sbp_op* op = new_op(ctx, SBP_OP_TABLE);
    int32_t rc = build_dense(op, N, G, h_D, N, h_w, 0);
    if (rc) {
        sbp_op_destroy(op);
        return rc;
    }
    *out = op;
    return SBP_OK;
}

int32_t sbp_op_sparse_create(sbp_ctx* ctx, int32_t N, const int32_t* h_rowptr, const int32_t* h_col,
                             const double* h_val, const double* h_w, double scale, sbp_op** out) {
    SBP_REQUIRE(ctx && h_rowptr && out, SBP_EINVAL, "null argument");
    SBP_REQUIRE(N >= 1, SBP_EDIM, "need N >= 1");
    SBP_REQUIRE(h_rowptr[0] == 0, SBP_EINVAL, "rowptr[0] must be 0");
    const int nnz = h_rowptr[N];
    SBP_REQUIRE(nnz == 0 || (h_col && h_val), SBP_EINVAL, "null col/val");
    for (int r = 0; r < N; r++) SBP_REQUIRE(h_rowptr[r + 1] >= h_rowptr[r], SBP_EINVAL, "rowptr not monotone");
    for (int j = 0; j < nnz; j++)
        SBP_REQUIRE(h_col[j] >= 0 && h_col[j] < N, SBP_EINVAL, "column index %d out of range at entry %d", h_col[j], j);
    SBP_CUDA(cudaSetDevice(ctx->device));
    sbp_op* op = new_op(ctx, SBP_OP_SPARSE);
    op->N = N;
    op->nnz = nnz;
    op->scale = scale;
    op->variant = SBP_KERNEL_CSR_SMEM;
    std::vector<int> rp(h_rowptr, h_rowptr + N + 1), col(h_col, h_col + nnz);
    std::vector<double> val(h_val, h_val + nnz);
    int32_t rc = build_sell(op, N, rp, col, val);
    if (!rc) rc = build_bsr(op, N, rp, col, val);
    if (!rc) rc = upload(&op->d_csr_rowptr, rp);
    if (!rc) rc = upload(&op->d_csr_col, col);
    if (!rc) rc = upload(&op->d_csr_val, val);
    for (int r = 0; r < N; r++) op->csr_max_row = std::max(op->csr_max_row, rp[r + 1] - rp[r]);
    if (!rc) {
        // kernel choice (option sparse_kernel / SBP_SPARSE_KERNEL=sell|bsr overrides): the tensor-core kernel whenever a tile fits in smem
        const bool want_bsr = ctx->tune[T_SPARSE_KERNEL] != 0;
        if (want_bsr && bsr_available(ctx, op)) op->variant = SBP_KERNEL_BSR_DMMA;
    }
    if (!rc && h_w) {
        std::vector<double> w(h_w, h_w + N);
        rc = upload(&op->d_w, w);
    }
    if (rc) {
        sbp_op_destroy(op);
        return rc;
    }
    *out = op;
    return SBP_OK;
}

int32_t sbp_op_sparse_from_dense(sbp_ctx* ctx, int32_t N, const double* h_D, int32_t ldd, const double* h_w,
                                 double scale, sbp_op** out) {
    SBP_REQUIRE(ctx && h_D && out, SBP_EINVAL, "null argument");
    SBP_REQUIRE(N >= 1 && ldd >= N, SBP_EDIM, "need N >= 1 and ldd >= N");
    std::vector<int> rp(N + 1, 0), col;
    std::vector<double> val;
    for (int i = 0; i < N; i++) {
        for (int k = 0; k < N; k++) {
            const double v = h_D[(size_t)k * ldd + i];
            if (v != 0.0) {
                col.push_back(k);
                val.push_back(v);
            }
        }
        rp[i + 1] = (int)col.size();
    }
    return sbp_op_sparse_create(ctx, N, rp.data(), col.data(), val.data(), h_w, scale, out);
}

int32_t sbp_op_advection2d_create(sbp_ctx* ctx, const sbp_op* A, const double* h_b, const double* h_w,
                                  const int32_t* h_mode, int32_t Nb, const int32_t* h_bnd_node,
                                  const double* h_bnd_tau, sbp_op** out) {
    SBP_REQUIRE(ctx && A && h_b && h_w && h_mode && out, SBP_EINVAL, "null argument");
    SBP_REQUIRE(A->kind == SBP_OP_DENSE || A->kind == SBP_OP_SPARSE, SBP_EINVAL, "A must be a dense or sparse operator");
    SBP_REQUIRE(Nb == 0 || (h_bnd_node && h_bnd_tau), SBP_EINVAL, "null boundary lists");
    const int N = A->N;
    for (int i = 0; i < N; i++) SBP_REQUIRE(h_mode[i] >= 0 && h_mode[i] <= 2, SBP_EINVAL, "mode[%d]=%d not in {0,1,2}", i, h_mode[i]);
    for (int i = 0; i < Nb; i++)
        SBP_REQUIRE(h_bnd_node[i] >= 0 && h_bnd_node[i] < N, SBP_EINVAL, "boundary node %d out of range", h_bnd_node[i]);
    SBP_CUDA(cudaSetDevice(ctx->device));
    sbp_op* op = new_op(ctx, SBP_OP_ADVECTION2D);
    op->N = N;
    op->nnz = A->nnz;
    op->inner = A;
    op->Nb = Nb;
    op->variant = A->variant;
    std::vector<double> b(h_b, h_b + N), qc(3 * (size_t)N, 0.0);
    std::vector<int> mode(h_mode, h_mode + N);
    for (int i = 0; i < N; i++) {
        qc[i] = h_w[i];
        qc[N + i] = h_w[i] * h_b[i];  // P*B diagonal (:90,:94-96)
    }
    for (int i = 0; i < Nb; i++) qc[2 * (size_t)N + h_bnd_node[i]] += h_bnd_tau[i];  // :98-105
    int32_t rc = upload(&op->d_b, b);
    if (!rc) rc = upload(&op->d_mode, mode);
    if (!rc && A->variant == SBP_KERNEL_BSR_DMMA) {
        // b_j of every block row in block order (quad A rows 0..3, quad B rows 4..7); 0 where no SAT applies
        const size_t nrb_pad = (size_t)((A->bsr_nrb + kBsrCH - 1) / kBsrCH) * kBsrCH;  // whole chunks: read without a guard
        std::vector<double> rbb(8 * nrb_pad, 0.0);
        for (int rb = 0; rb < A->bsr_nrb; rb++)
            for (int r = 0; r < 8; r++) {
                const int i = (r < 4 ? A->bsr_rows[2 * rb] : A->bsr_rows[2 * rb + 1] - 4) + r;
                if (i < A->bsr_N && h_mode[i % N] == 1) rbb[8 * (size_t)rb + r] = h_b[i % N];  // bsr_N = 2N: instance pairs
            }
        rc = upload(&op->d_rb_bdiag, rbb);
        // fragments of A - diag(b): the SAT's u-term rides on the tensor pipe (bsr3.cu, satfold)
        if (!rc && A->bsr_frags > 0 && A->bsr_diag.size() == 8 * (size_t)A->bsr_nrb) {
            const size_t nv = (size_t)A->bsr_frags * 32;
            std::vector<double> img(nv);
            std::vector<unsigned char> mask(nv, 0);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
                cudaMemcpy(img.data(), A->d_bsr_val, nv * 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
                set_error("advection2d_create: device copy failed");
                rc = SBP_ECUDA;
            }
            for (int rb = 0; rb < A->bsr_nrb && !rc; rb++)
                for (int r = 0; r < 8; r++) {
                    const int k = A->bsr_diag[8 * (size_t)rb + r];
                    if (k < 0) continue;
                    mask[k] = 1;
                    img[k] -= rbb[8 * (size_t)rb + r] / A->scale;  // launch_bsr multiplies the image by alpha * scale
                }
            if (!rc) rc = upload(&op->d_bsr_val_sat, img);
            if (!rc) rc = upload(&op->d_bsr_diagmask, mask);
            if (!rc && cudaMalloc(&op->d_bsr_val_stage, nv * 8) != cudaSuccess) rc = SBP_ENOMEM;
        }
    }
    if (!rc) rc = upload(&op->d_q_consts, qc);
    if (!rc) {
        std::vector<double> w(h_w, h_w + N);
        rc = upload(&op->d_w, w);
    }
    if (rc) {
        sbp_op_destroy(op);
        return rc;
    }
    *out = op;
    return SBP_OK;
}

int32_t sbp_op_advection1d_create(sbp_ctx* ctx, const sbp_op* D, const double* h_a, sbp_op** out) {
    SBP_REQUIRE(ctx && D && h_a && out, SBP_EINVAL, "null argument");
    SBP_REQUIRE(D->kind == SBP_OP_DENSE || D->kind == SBP_OP_SPARSE, SBP_EINVAL, "D must be a dense or sparse operator");
    SBP_REQUIRE(D->d_w != nullptr, SBP_EINVAL, "the 1D advection bundle needs an operator with weights");
    const int N = D->N;
    SBP_CUDA(cudaSetDevice(ctx->device));
    sbp_op* op = new_op(ctx, SBP_OP_ADVECTION1D);
    op->N = N;
    op->nnz = D->nnz;
    op->inner = D;
    op->variant = D->variant;
    bool unit = true;
    for (int i = 0; i < N; i++) unit = unit && (h_a[i] == 1.0);
    // D*a on the device (one instance), then w .* (D*a) on the host
    std::vector<double> a(h_a, h_a + N), w(N), Da(N);
    double *d_a = nullptr, *d_Da = nullptr;
    int32_t rc = upload(&d_a, a);
    if (rc) goto fail;
    if (cudaMalloc(&d_Da, N * sizeof(double)) != cudaSuccess) {
        rc = SBP_ENOMEM;
        goto fail;
    }
    rc = sbp_apply(ctx, D, d_Da, d_a, 1, 1.0, 0.0);
    if (rc) goto fail;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
        cudaMemcpy(Da.data(), d_Da, N * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(w.data(), D->d_w, N * 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("advection1d_create: device copy failed");
        rc = SBP_ECUDA;
        goto fail;
    }
    {
        std::vector<double> qc(3 * (size_t)N);
        for (int i = 0; i < N; i++) {
            qc[i] = w[i];
            qc[N + i] = w[i] * Da[i];
            qc[2 * (size_t)N + i] = h_a[i];
        }
        op->sat_a0 = h_a[0], op->sat_aN = h_a[N - 1], op->sat_w0 = w[0], op->sat_wN = w[N - 1];
        rc = upload(&op->d_q_consts, qc);
        if (rc) goto fail;
    }
    if (!unit) {
        if (D->kind == SBP_OP_DENSE && is_small(D)) {
            const int KS = 2 * D->NT;
            std::vector<double> cf((size_t)KS * 4, 0.0);
            for (int ks = 0; ks < KS; ks++)
                for (int t = 0; t < 4; t++) {
                    const int k = frag_k(D->layout, ks, t);
                    if (k < N) cf[ks * 4 + t] = h_a[k];
                }
            rc = upload(&op->d_a, cf);
        } else if (D->kind == SBP_OP_SPARSE) {
            rc = upload(&op->d_a, a);
        } else {
            set_error("advection1d: variable speed with a large dense operator is not supported");
            rc = SBP_EUNSUPPORTED;
        }
        if (rc) goto fail;
    }
    cudaFree(d_a);
    cudaFree(d_Da);
    *out = op;
    return SBP_OK;
fail:
    if (d_a) cudaFree(d_a);
    if (d_Da) cudaFree(d_Da);
    sbp_op_destroy(op);
    return rc;
}

int32_t sbp_op_info(const sbp_op* op, int32_t* kind, int32_t* N, int32_t* G, int64_t* nnz, int32_t* variant) {
    SBP_REQUIRE(op, SBP_EINVAL, "null op");
    if (kind) *kind = op->kind;
    if (N) *N = op->N;
    if (G) *G = op->G;
    if (nnz) *nnz = op->nnz;
    if (variant) *variant = op->variant;
    return SBP_OK;
}

// ---- hot path -----------------------------------------------------------------------------------
// tuning knob of the even/odd kernel (bit0 prefetch, bit1 one m-tile per warp, bit2 256-thread CTAs);
// the default was picked from the sweep in profiles/ (tools/tune_sym.py)
static int sym_variant(const sbp_ctx* ctx) { return ctx->tune[T_SYM_VARIANT] & 7; }
static int32_t check_batch(sbp_ctx* ctx, const sbp_op* op, const void* dU, const void* U, int64_t B) {
    SBP_REQUIRE(ctx && op, SBP_EINVAL, "null ctx/op");
    SBP_REQUIRE(op->ctx == ctx, SBP_EINVAL, "operator belongs to a different context");
    SBP_REQUIRE(B >= 0, SBP_EDIM, "negative batch size");
    SBP_REQUIRE(B == 0 || (dU && U), SBP_EINVAL, "null batch pointer");
    SBP_REQUIRE(B == 0 || dU != U, SBP_EINVAL, "dU and U must not alias");
    SBP_REQUIRE((reinterpret_cast<uintptr_t>(dU) & 7) == 0 && (reinterpret_cast<uintptr_t>(U) & 7) == 0, SBP_EINVAL,
                "batch pointers must be 8-byte aligned");
    return SBP_OK;
}

int32_t sbp_apply(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                  double beta) {
    int32_t rc = check_batch(ctx, op, dU, U, B);
    if (rc || B == 0) return rc;
    SBP_CUDA(cudaSetDevice(ctx->device));
    switch (op->kind) {
        case SBP_OP_DENSE:
            if (op->sym) return launch_dense_small_sym(ctx, op, dU, U, B, alpha, beta, nullptr, nullptr, sym_variant(ctx));
            if (is_small(op)) return launch_dense_small(ctx, op, dU, U, B, alpha, beta, false, nullptr, nullptr);
            return launch_dense_gemm(ctx, op, dU, U, B, alpha, beta);
        case SBP_OP_TABLE:
            if (is_small(op)) return launch_dense_small(ctx, op, dU, U, B, alpha, beta, true, nullptr, nullptr);
            return launch_table_generic(ctx, op, dU, U, B, nullptr, alpha, beta, nullptr, nullptr);
        case SBP_OP_SPARSE:
            if (bsr_can_run(op, U)) return launch_bsr(ctx, op, dU, U, B, alpha, beta, nullptr, nullptr, 0, nullptr);
            return launch_sparse(ctx, op, dU, U, B, alpha, beta, nullptr, nullptr, 0, nullptr);
        default:
            set_error("sbp_apply: operator kind %d is a semidiscretization bundle; use sbp_rhs_*", op->kind);
            return SBP_EINVAL;
    }
}

int32_t sbp_apply_table(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B,
                        const int32_t* op_index, double alpha, double beta, double* energy, double* energy_sum) {
    int32_t rc = check_batch(ctx, op, dU, U, B);
    if (rc) return rc;
    SBP_REQUIRE(op->kind == SBP_OP_TABLE || op->kind == SBP_OP_DENSE, SBP_EINVAL,
                "sbp_apply_table needs a table or dense operator");
    SBP_CUDA(cudaSetDevice(ctx->device));
    if (B == 0) {
        if (energy_sum) SBP_CUDA(cudaMemsetAsync(energy_sum, 0, sizeof(double), ctx->stream));
        return SBP_OK;
    }
    if (op->variant == SBP_KERNEL_DGEMM_TILED) {
        rc = launch_dense_gemm(ctx, op, dU, U, B, alpha, beta);
        if (rc || !(energy || energy_sum)) return rc;
        SBP_REQUIRE(op->d_w, SBP_EINVAL, "energy requested but the operator has no weights");
        return launch_integrate(ctx, op->d_w, op->N, 1, U, nullptr, B, 1, energy, energy_sum);
    }
    if (op_index == nullptr && op->sym)
        return launch_dense_small_sym(ctx, op, dU, U, B, alpha, beta, energy, energy_sum, sym_variant(ctx));
    if (op_index == nullptr && is_small(op))
        return launch_dense_small(ctx, op, dU, U, B, alpha, beta, true, energy, energy_sum);
    return launch_table_generic(ctx, op, dU, U, B, op_index, alpha, beta, energy, energy_sum);
}

int32_t sbp_integrate(sbp_ctx* ctx, const sbp_op* op, const double* U, const double* V, int64_t B, int32_t mode,
                      double* out, double* out_sum) {
    SBP_REQUIRE(ctx && op, SBP_EINVAL, "null ctx/op");
    SBP_REQUIRE(op->d_w, SBP_EINVAL, "operator has no quadrature weights");
    SBP_REQUIRE(B >= 0, SBP_EDIM, "negative batch size");
    SBP_CUDA(cudaSetDevice(ctx->device));
    if (B == 0) {
        if (out_sum) SBP_CUDA(cudaMemsetAsync(out_sum, 0, sizeof(double), ctx->stream));
        return SBP_OK;
    }
    SBP_REQUIRE(U, SBP_EINVAL, "null U");
    return launch_integrate(ctx, op->d_w, op->N, op->G, U, V, B, mode, out, out_sum);
}

int32_t sbp_scale_by_mass_matrix(sbp_ctx* ctx, const sbp_op* op, double* U, int64_t B, double factor,
                                 int32_t inverse) {
    SBP_REQUIRE(ctx && op, SBP_EINVAL, "null ctx/op");
    SBP_REQUIRE(op->d_w, SBP_EINVAL, "operator has no quadrature weights");
    SBP_REQUIRE(op->G == 1, SBP_EUNSUPPORTED, "scale_by_mass_matrix on operator tables is not supported");
    SBP_REQUIRE(B >= 0, SBP_EDIM, "negative batch size");
    if (B == 0) return SBP_OK;
    SBP_REQUIRE(U, SBP_EINVAL, "null U");
    SBP_CUDA(cudaSetDevice(ctx->device));
    return launch_scale(ctx, op->d_w, op->N, U, B, factor, inverse);
}

int32_t sbp_rhs_advection2d(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, const double* g,
                            int64_t g_stride, double* quantities) {
    int32_t rc = check_batch(ctx, op, dU, U, B);
    if (rc || B == 0) return rc;
    SBP_REQUIRE(op->kind == SBP_OP_ADVECTION2D, SBP_EINVAL, "not a 2D advection bundle");
    SBP_REQUIRE(g != nullptr, SBP_EINVAL, "boundary data g is required");
    SBP_REQUIRE(g_stride == 0 || g_stride >= op->N, SBP_EDIM, "g_stride must be 0 or >= N");
    SBP_CUDA(cudaSetDevice(ctx->device));
    const sbp_op* A = op->inner;
    if (A->kind == SBP_OP_SPARSE && bsr_can_run(A, U)) {
        rc = launch_bsr(ctx, A, dU, U, B, -1.0, 0.0, op, g, g_stride, nullptr);
        if (rc || !quantities) return rc;
        return launch_adv2d_epilogue(ctx, op, dU, U, B, g, g_stride, quantities, false);
    }
    if (A->kind == SBP_OP_SPARSE) return launch_sparse(ctx, A, dU, U, B, -1.0, 0.0, op, g, g_stride, quantities);
    rc = sbp_apply(ctx, A, dU, U, B, -1.0, 0.0);
    if (rc) return rc;
    return launch_adv2d_epilogue(ctx, op, dU, U, B, g, g_stride, quantities, true);
}

int32_t sbp_rhs_advection1d(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B,
                            const double* bc, int64_t bc_stride, double* quantities) {
    int32_t rc = check_batch(ctx, op, dU, U, B);
    if (rc || B == 0) return rc;
    SBP_REQUIRE(op->kind == SBP_OP_ADVECTION1D, SBP_EINVAL, "not a 1D advection bundle");
    SBP_REQUIRE(bc != nullptr, SBP_EINVAL, "boundary data bc is required");
    SBP_REQUIRE(bc_stride == 0 || bc_stride >= 2, SBP_EDIM, "bc_stride must be 0 or >= 2");
    SBP_CUDA(cudaSetDevice(ctx->device));
    const sbp_op* D = op->inner;
    // the SATs at the two end nodes are fused into the apply kernel (one pass, no second launch)
    Sat1D sat;
    sat.bc = bc;
    sat.bc_stride = bc_stride;
    sat.a0 = op->sat_a0, sat.aN = op->sat_aN, sat.w0 = op->sat_w0, sat.wN = op->sat_wN;
    sat.enabled = 1;
    bool fused = true;
    if (D->kind == SBP_OP_SPARSE && bsr_can_run(D, U) && op->d_a == nullptr)
        rc = launch_bsr(ctx, D, dU, U, B, -1.0, 0.0, nullptr, nullptr, 0, &sat);
    else if (D->kind == SBP_OP_SPARSE)
        rc = launch_sparse_ex(ctx, D, dU, U, B, -1.0, 0.0, nullptr, nullptr, 0, op->d_a, &sat);
    else if (D->sym && op->d_a == nullptr)
        rc = launch_dense_small_sym(ctx, D, dU, U, B, -1.0, 0.0, nullptr, nullptr, sym_variant(ctx), &sat);
    else if (is_small(D))
        rc = launch_dense_small_ex(ctx, D, dU, U, B, -1.0, 0.0, nullptr, nullptr, op->d_a, &sat);
    else {
        rc = launch_dense_gemm(ctx, D, dU, U, B, -1.0, 0.0);
        fused = false;
    }
    if (rc) return rc;
    if (fused && quantities == nullptr) return SBP_OK;
    return launch_adv1d_epilogue(ctx, op, dU, U, B, bc, bc_stride, quantities, !fused);
}

// ---- Runge-Kutta stage updates fused into the RHS pass ----------------------------------------------
// Unext <- ca * U + cb * Uextra + ch * RHS(U).  On the block-sparse kernel the update rides on the RHS pass (U's own rows are
// in the box ring, Uextra is read by the lane that writes the same entries); every other kernel family evaluates the RHS
// into scratch and combines in one pass (sbp_lincomb).
static int32_t stage_fallback(sbp_ctx* ctx, double* Unext, const double* U, int64_t n, const double* Uextra, double ca,
                              double cb, double ch, const double* rhs) {
    return launch_lincomb(ctx, Unext, U, Uextra, rhs, n, ca, 0.0, Uextra ? cb : 0.0, ch);
}
static int32_t ensure_stage_buf(sbp_ctx* ctx, size_t doubles) {
    if (ctx->stage_doubles >= doubles) return SBP_OK;
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->stage_buf) cudaFree(ctx->stage_buf);
    ctx->stage_buf = nullptr, ctx->stage_doubles = 0;
    SBP_CUDA(cudaMalloc(&ctx->stage_buf, doubles * sizeof(double)));
    ctx->stage_doubles = doubles;
    return SBP_OK;
}
static bool stage_fusion_enabled(const sbp_ctx* ctx) { return ctx->tune[T_FUSED_STAGE] != 0; }  // 0: always RHS + sbp_lincomb

int32_t sbp_rhs_advection2d_stage(sbp_ctx* ctx, const sbp_op* op, double* Unext, const double* U, int64_t B,
                                  const double* g, int64_t g_stride, double ca, double cb, const double* Uextra,
                                  double ch) {
    int32_t rc = check_batch(ctx, op, Unext, U, B);
    if (rc || B == 0) return rc;
    SBP_REQUIRE(op->kind == SBP_OP_ADVECTION2D, SBP_EINVAL, "not a 2D advection bundle");
    SBP_REQUIRE(g != nullptr, SBP_EINVAL, "boundary data g is required");
    SBP_REQUIRE(g_stride == 0 || g_stride >= op->N, SBP_EDIM, "g_stride must be 0 or >= N");
    SBP_REQUIRE(Unext != U, SBP_EINVAL, "Unext must not alias U (other instances' rows of U are still being read)");
    SBP_CUDA(cudaSetDevice(ctx->device));
    const sbp_op* A = op->inner;
    const bool al = ((reinterpret_cast<uintptr_t>(Uextra) | reinterpret_cast<uintptr_t>(Unext)) & 15) == 0;
    if (stage_fusion_enabled(ctx) && A->kind == SBP_OP_SPARSE && bsr_can_run(A, U) && al) {
        const StageUpdate su{ca, cb, ch, Uextra};
        return launch_bsr(ctx, A, Unext, U, B, -1.0, 0.0, op, g, g_stride, nullptr, &su);
    }
    const int64_t n = B * (int64_t)op->N;
    rc = ensure_stage_buf(ctx, (size_t)n);
    if (rc) return rc;
    double* rhs = ctx->stage_buf;
    rc = sbp_rhs_advection2d(ctx, op, rhs, U, B, g, g_stride, nullptr);
    if (rc) return rc;
    return stage_fallback(ctx, Unext, U, n, Uextra, ca, cb, ch, rhs);
}

int32_t sbp_rhs_advection1d_stage(sbp_ctx* ctx, const sbp_op* op, double* Unext, const double* U, int64_t B,
                                  const double* bc, int64_t bc_stride, double ca, double cb, const double* Uextra,
                                  double ch) {
    int32_t rc = check_batch(ctx, op, Unext, U, B);
    if (rc || B == 0) return rc;
    SBP_REQUIRE(op->kind == SBP_OP_ADVECTION1D, SBP_EINVAL, "not a 1D advection bundle");
    SBP_REQUIRE(bc != nullptr, SBP_EINVAL, "boundary data bc is required");
    SBP_REQUIRE(bc_stride == 0 || bc_stride >= 2, SBP_EDIM, "bc_stride must be 0 or >= 2");
    SBP_REQUIRE(Unext != U, SBP_EINVAL, "Unext must not alias U (other instances' rows of U are still being read)");
    SBP_CUDA(cudaSetDevice(ctx->device));
    const sbp_op* D = op->inner;
    const bool al = ((reinterpret_cast<uintptr_t>(Uextra) | reinterpret_cast<uintptr_t>(Unext)) & 15) == 0;
    if (stage_fusion_enabled(ctx) && D->kind == SBP_OP_SPARSE && bsr_can_run(D, U) && op->d_a == nullptr && al) {
        Sat1D sat;
        sat.bc = bc;
        sat.bc_stride = bc_stride;
        sat.a0 = op->sat_a0, sat.aN = op->sat_aN, sat.w0 = op->sat_w0, sat.wN = op->sat_wN;
        sat.enabled = 1;
        const StageUpdate su{ca, cb, ch, Uextra};
        return launch_bsr(ctx, D, Unext, U, B, -1.0, 0.0, nullptr, nullptr, 0, &sat, &su);
    }
    const int64_t n = B * (int64_t)op->N;
    rc = ensure_stage_buf(ctx, (size_t)n);
    if (rc) return rc;
    double* rhs = ctx->stage_buf;
    rc = sbp_rhs_advection1d(ctx, op, rhs, U, B, bc, bc_stride, nullptr);
    if (rc) return rc;
    return stage_fallback(ctx, Unext, U, n, Uextra, ca, cb, ch, rhs);
}

int32_t sbp_axpbypcz(sbp_ctx* ctx, double* Y, const double* X, const double* Z, int64_t n, double a, double b,
                     double c) {
    SBP_REQUIRE(ctx, SBP_EINVAL, "null ctx");
    SBP_REQUIRE(n >= 0, SBP_EDIM, "negative length");
    if (n == 0) return SBP_OK;
    SBP_REQUIRE(Y && X, SBP_EINVAL, "null vector");
    SBP_CUDA(cudaSetDevice(ctx->device));
    return launch_axpbypcz(ctx, Y, X, Z, n, a, b, c);
}

int32_t sbp_lincomb(sbp_ctx* ctx, double* Y, const double* X, const double* Z, const double* W, int64_t n, double a,
                    double b, double c, double d) {
    SBP_REQUIRE(ctx, SBP_EINVAL, "null ctx");
    SBP_REQUIRE(n >= 0, SBP_EDIM, "negative length");
    if (n == 0) return SBP_OK;
    SBP_REQUIRE(Y && X, SBP_EINVAL, "null vector");
    SBP_CUDA(cudaSetDevice(ctx->device));
    return launch_lincomb(ctx, Y, X, Z, W, n, a, b, c, d);
}

// ---- multi-GPU --------------------------------------------------------------------------------------
int32_t sbp_comm_unique_id(char id[SBP_NCCL_ID_BYTES]) {
    SBP_REQUIRE(id, SBP_EINVAL, "null id");
    int32_t rc = load_nccl();
    if (rc) return rc;
    int e = g_nccl.GetUniqueId(id);
    SBP_REQUIRE(e == 0, SBP_ENCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error");
    return SBP_OK;
}
int32_t sbp_comm_init(sbp_ctx* ctx, int32_t nranks, int32_t rank, const char id[SBP_NCCL_ID_BYTES]) {
    SBP_REQUIRE(ctx && id, SBP_EINVAL, "null argument");
    SBP_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, SBP_EINVAL, "bad rank %d / %d", rank, nranks);
    int32_t rc = load_nccl();
    if (rc) return rc;
    SBP_CUDA(cudaSetDevice(ctx->device));
    Id128 uid;
    memcpy(uid.b, id, 128);
    int e = g_nccl.CommInitRank(&ctx->nccl_comm, nranks, uid, rank);
    SBP_REQUIRE(e == 0, SBP_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error");
    return SBP_OK;
}
int32_t sbp_allreduce_sum(sbp_ctx* ctx, double* dptr, int32_t count) {
    SBP_REQUIRE(ctx && dptr, SBP_EINVAL, "null argument");
    SBP_REQUIRE(ctx->nccl_comm, SBP_ENCCL, "communicator not initialised (sbp_comm_init)");
    int e = g_nccl.AllReduce(dptr, dptr, (size_t)count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, ctx->nccl_comm, ctx->stream);
    SBP_REQUIRE(e == 0, SBP_ENCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error");
    return SBP_OK;
}
int32_t sbp_comm_destroy(sbp_ctx* ctx) {
    SBP_REQUIRE(ctx, SBP_EINVAL, "null ctx");
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    return SBP_OK;
}

// ---- host-buffer entry points ---------------------------------------------------------------------------
int32_t sbp_apply_host(sbp_ctx* ctx, const sbp_op* op, double* h_dU, const double* h_U, int64_t B, double alpha,
                       double beta, int64_t chunk_instances) {
    SBP_REQUIRE(ctx && op, SBP_EINVAL, "null ctx/op");
    SBP_REQUIRE(B >= 0, SBP_EDIM, "negative batch size");
    if (B == 0) return SBP_OK;
    SBP_REQUIRE(h_dU && h_U, SBP_EINVAL, "null host buffer");
    SBP_CUDA(cudaSetDevice(ctx->device));
    return host_pipeline(ctx, op->N, op->G, h_dU, h_U, B, beta != 0.0, chunk_instances,
                         [&](double* dOut, const double* dIn, int64_t nb, int64_t) {
                             return sbp_apply(ctx, op, dOut, dIn, nb, alpha, beta);
                         });
}

int32_t sbp_rhs_advection2d_host(sbp_ctx* ctx, const sbp_op* op, double* h_dU, const double* h_U, int64_t B,
                                 const double* h_g, int64_t g_stride, int64_t chunk_instances) {
    SBP_REQUIRE(ctx && op, SBP_EINVAL, "null ctx/op");
    SBP_REQUIRE(op->kind == SBP_OP_ADVECTION2D, SBP_EINVAL, "not a 2D advection bundle");
    SBP_REQUIRE(B >= 0, SBP_EDIM, "negative batch size");
    if (B == 0) return SBP_OK;
    SBP_REQUIRE(h_dU && h_U && h_g, SBP_EINVAL, "null host buffer");
    SBP_REQUIRE(g_stride == 0 || g_stride >= op->N, SBP_EDIM, "g_stride must be 0 or >= N");
    SBP_CUDA(cudaSetDevice(ctx->device));
    const int N = op->N;
    // boundary data: shared by the batch (N doubles) or per instance (B rows of g_stride doubles), uploaded once up front
    const size_t gbytes = (g_stride == 0 ? (size_t)N : (size_t)((B - 1) * g_stride + N)) * sizeof(double);
    void* d_g = nullptr;
    int32_t rc = ensure_aux(ctx, 0, gbytes, &d_g);
    if (rc) return rc;
    SBP_CUDA(cudaMemcpyAsync(d_g, h_g, gbytes, cudaMemcpyHostToDevice, ctx->stream));
    return host_pipeline(ctx, N, 1, h_dU, h_U, B, false, chunk_instances,
                         [&](double* dOut, const double* dIn, int64_t nb, int64_t done) {
                             return sbp_rhs_advection2d(ctx, op, dOut, dIn, nb,
                                                        static_cast<const double*>(d_g) + done * g_stride, g_stride, nullptr);
                         });
}

// ---- device-resident fixed-step SSPRK53 ---------------------------------------------------------------
// Shu-Osher form of OrdinaryDiffEq's SSPRK53 (Ruuth 2006), the integrator of examples/RBF_FSBP_advection.jl:45-51
static const double kSsp[] = {/*a30*/ 0.355909775063327, /*a32*/ 0.644090224936674, /*a40*/ 0.367933791638137,
                              /*a43*/ 0.632066208361863, /*a52*/ 0.237593836598569, /*a54*/ 0.762406163401431,
                              /*b10*/ 0.377268915331368, /*b21*/ 0.377268915331368, /*b32*/ 0.242995220537396,
                              /*b43*/ 0.238458932846290, /*b54*/ 0.287632146308408};

int32_t sbp_solve_ssprk53(sbp_ctx* ctx, const sbp_op* op, double* U, int64_t B, const double* bc_table,
                          int64_t bc_row_stride, int64_t nsteps, const double* h_steps, double* work) {
    SBP_REQUIRE(ctx && op, SBP_EINVAL, "null ctx/op");
    SBP_REQUIRE(op->kind == SBP_OP_ADVECTION2D || op->kind == SBP_OP_ADVECTION1D, SBP_EINVAL,
                "sbp_solve_ssprk53 needs an advection bundle");
    SBP_REQUIRE(B >= 0 && nsteps >= 0, SBP_EDIM, "negative batch size / step count");
    if (B == 0 || nsteps == 0) return SBP_OK;
    SBP_REQUIRE(U && bc_table && h_steps, SBP_EINVAL, "null argument");
    const bool two_d = op->kind == SBP_OP_ADVECTION2D;
    SBP_REQUIRE(bc_row_stride >= (two_d ? op->N : 2), SBP_EDIM, "bc_row_stride too small (%lld)", (long long)bc_row_stride);
    SBP_CUDA(cudaSetDevice(ctx->device));
    // small batches: the whole solve in ONE launch (a CTA per instance, state in shared memory; solve_small.cu) instead of
    // 5 launches per step.  Option solve_small (SBP_SOLVE_SMALL) = 0 forces the batched path, 1 the single launch whenever it is applicable.
    {
        const int force = ctx->tune[T_SOLVE_SMALL];
        const bool applicable = solve_small_applies(ctx, op, force == 1 ? 1 : B);
        if (force != 0 && applicable) {
            void* d_h = nullptr;
            int32_t rc = ensure_aux(ctx, 3, (size_t)nsteps * sizeof(double), &d_h);
            if (rc) return rc;
            SBP_CUDA(cudaMemcpyAsync(d_h, h_steps, (size_t)nsteps * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            return launch_solve_small(ctx, op, U, B, bc_table, bc_row_stride, nsteps, static_cast<const double*>(d_h));
        }
    }
    const size_t n = (size_t)B * op->N;
    if (!work) {
        void* w = nullptr;
        int32_t rc = ensure_aux(ctx, 1, 3 * n * sizeof(double), &w);
        if (rc) return rc;
        work = static_cast<double*>(w);
    }
    double *u = U, *u1 = work, *u2 = work + n, *u3 = work + 2 * n;
    auto stage = [&](double* unext, const double* uin, const double* bc, double ca, double ch, double cb, const double* extra) {
        return two_d ? sbp_rhs_advection2d_stage(ctx, op, unext, uin, B, bc, 0, ca, cb, extra, ch)
                     : sbp_rhs_advection1d_stage(ctx, op, unext, uin, B, bc, 0, ca, cb, extra, ch);
    };
    const double* k = kSsp;
    for (int64_t s = 0; s < nsteps; s++) {
        const double h = h_steps[s];
        const double* bc = bc_table + 5 * s * bc_row_stride;
        int32_t rc = stage(u1, u, bc, 1.0, k[6] * h, 0.0, nullptr);                                   // u1 = u + b10 h f(u)
        if (!rc) rc = stage(u2, u1, bc + bc_row_stride, 1.0, k[7] * h, 0.0, nullptr);                 // u2 = u1 + b21 h f(u1)
        if (!rc) rc = stage(u3, u2, bc + 2 * bc_row_stride, k[1], k[8] * h, k[0], u);                 // u3 = a30 u + a32 u2 + b32 h f(u2)
        if (!rc) rc = stage(u1, u3, bc + 3 * bc_row_stride, k[3], k[9] * h, k[2], u);                 // u4 = a40 u + a43 u3 + b43 h f(u3)
        if (!rc) rc = stage(u, u1, bc + 4 * bc_row_stride, k[5], k[10] * h, k[4], u2);                // u  = a52 u2 + a54 u4 + b54 h f(u4)
        if (rc) return rc;
    }
    return SBP_OK;
}

int32_t sbp_solve_ssprk53_host(sbp_ctx* ctx, const sbp_op* op, double* h_U, int64_t B, const double* h_bc_table,
                               int64_t bc_row_stride, int64_t nsteps, const double* h_steps) {
    SBP_REQUIRE(ctx && op, SBP_EINVAL, "null ctx/op");
    SBP_REQUIRE(B >= 0 && nsteps >= 0, SBP_EDIM, "negative batch size / step count");
    if (B == 0) return SBP_OK;
    SBP_REQUIRE(h_U && (nsteps == 0 || (h_bc_table && h_steps)), SBP_EINVAL, "null argument");
    SBP_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)B * op->N, tbytes = (size_t)(5 * nsteps * bc_row_stride) * sizeof(double);
    void *d_u = nullptr, *d_t = nullptr;
    int32_t rc = ensure_aux(ctx, 2, n * sizeof(double), &d_u);
    if (!rc) rc = ensure_aux(ctx, 0, std::max<size_t>(tbytes, 16), &d_t);
    if (rc) return rc;
    SBP_CUDA(cudaMemcpyAsync(d_u, h_U, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (tbytes) SBP_CUDA(cudaMemcpyAsync(d_t, h_bc_table, tbytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = sbp_solve_ssprk53(ctx, op, static_cast<double*>(d_u), B, static_cast<const double*>(d_t), bc_row_stride, nsteps,
                           h_steps, nullptr);
    if (rc) return rc;
    SBP_CUDA(cudaMemcpyAsync(h_U, d_u, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    return SBP_OK;
}

// ---- PCIe / host-memory copy ceiling of this box (what bounds the *_host entry points) ----------------------------------
int32_t sbp_pcie_probe(sbp_ctx* ctx, size_t bytes, double* h2d_gbs, double* d2h_gbs, double* bidir_gbs) {
    SBP_REQUIRE(ctx && bytes >= (1u << 20), SBP_EINVAL, "sbp_pcie_probe: need a ctx and at least 1 MiB");
    SBP_CUDA(cudaSetDevice(ctx->device));
    void *h_a = nullptr, *h_b = nullptr, *d_a = nullptr, *d_b = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    int32_t rc = SBP_OK;
    auto fail = [&](const char* what) {
        set_error("sbp_pcie_probe: %s failed: %s", what, cudaGetErrorString(cudaGetLastError()));
        rc = SBP_ECUDA;
    };
    if (cudaHostAlloc(&h_a, bytes, cudaHostAllocDefault) != cudaSuccess || cudaHostAlloc(&h_b, bytes, cudaHostAllocDefault) != cudaSuccess ||
        cudaMalloc(&d_a, bytes) != cudaSuccess || cudaMalloc(&d_b, bytes) != cudaSuccess)
        fail("allocation");
    if (!rc) {
        memset(h_a, 1, bytes), memset(h_b, 2, bytes);
        cudaEventCreate(&e0), cudaEventCreate(&e1), cudaEventCreate(&e2);
        auto timed = [&](bool up, bool down) {
            double best = 0.0;
            for (int rep = 0; rep < 3; rep++) {
                cudaStreamSynchronize(ctx->s_in), cudaStreamSynchronize(ctx->s_out);
                cudaEventRecord(e0, ctx->s_in);
                cudaStreamWaitEvent(ctx->s_out, e0, 0);
                if (up) cudaMemcpyAsync(d_a, h_a, bytes, cudaMemcpyHostToDevice, ctx->s_in);
                if (down) cudaMemcpyAsync(h_b, d_b, bytes, cudaMemcpyDeviceToHost, ctx->s_out);
                cudaEventRecord(e2, ctx->s_out);
                cudaStreamWaitEvent(ctx->s_in, e2, 0);
                cudaEventRecord(e1, ctx->s_in);
                if (cudaEventSynchronize(e1) != cudaSuccess) return -1.0;
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                best = std::max(best, (double)bytes * ((up ? 1 : 0) + (down ? 1 : 0)) / (ms * 1e6));
            }
            return best;
        };
        const double u = timed(true, false), d = timed(false, true), b = timed(true, true);
        if (u < 0 || d < 0 || b < 0) fail("timed copy");
        if (h2d_gbs) *h2d_gbs = u;
        if (d2h_gbs) *d2h_gbs = d;
        if (bidir_gbs) *bidir_gbs = b;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e2) cudaEventDestroy(e2);
    if (h_a) cudaFreeHost(h_a);
    if (h_b) cudaFreeHost(h_b);
    if (d_a) cudaFree(d_a);
    if (d_b) cudaFree(d_b);
    return rc;
}

}  // extern "C"
