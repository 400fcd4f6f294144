/* sbp_b200.h -- C ABI of libsbp_b200.so (CUDA sm_100a).
 *
 * Drop-in boundary for the data-parallel hot path of SummationByPartsOperatorsExtra.jl:
 * batched application of assembled SBP derivative operators, the fused linear-advection RHS
 * with SAT boundary terms, and the P-weighted reductions (mass / energy / analysis quantities).
 * The reference has no FFI layer (pure Julia multiple dispatch); each entry point below names the
 * reference method it replaces (paths relative to the reference repository root).  The Julia shim
 * that binds these with `ccall` is shown in INTEGRATION.md and shipped as
 * summationbypartsoperatorsextra.jl_b200/julia/SBPB200.jl.
 *
 * Conventions
 *  - every function returns an int32 status: 0 = SBP_OK, < 0 = error; the message of the last
 *    error on the calling thread is returned by sbp_last_error().
 *  - "batch" layout: a batch of B nodal vectors of N nodes is ONE array of B*N doubles with each
 *    instance contiguous (Julia `Matrix{Float64}(N, B)`, column-major; numpy (B, N) C-order).
 *    Instance b lives at U + b*N.
 *  - pointers named U, dU, V, g, out, ... are DEVICE pointers (from sbp_malloc or any CUDA
 *    allocation) unless the name starts with h_.  Host arrays passed to *_create are copied;
 *    they are borrowed only for the duration of the call.
 *  - one sbp_ctx per device and host thread; calls on a ctx are stream-ordered and asynchronous;
 *    sbp_ctx_sync / sbp_memcpy_d2h synchronise.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    SBP_ECUDA.
 */
#ifndef SBP_B200_H
#define SBP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBP_OK 0
#define SBP_EINVAL -1   /* bad argument (maps to Julia ArgumentError)                       */
#define SBP_EDIM -2     /* size mismatch (maps to DimensionMismatch / @argcheck ArgumentError,
                           src/polynomialbases_operators.jl:143-146, src/subcell_operators.jl:195) */
#define SBP_ECUDA -3    /* CUDA runtime / launch failure, or no device                      */
#define SBP_ENOMEM -4   /* device or host allocation failed                                 */
#define SBP_ENCCL -5    /* NCCL unavailable or failed                                       */
#define SBP_EUNSUPPORTED -6

typedef struct sbp_ctx sbp_ctx; /* device + stream + scratch                                  */
typedef struct sbp_op sbp_op;   /* immutable device-resident operator                         */

/* operator kinds (sbp_op_info) */
#define SBP_OP_DENSE 1
#define SBP_OP_SPARSE 2
#define SBP_OP_TABLE 3
#define SBP_OP_ADVECTION2D 4
#define SBP_OP_ADVECTION1D 5

/* flags for sbp_op_dense_create */
#define SBP_DENSE_AUTO 0        /* pick kernel from N (smem-resident DMMA tiles or tiled DGEMM)  */
#define SBP_DENSE_FORCE_GEMM 1  /* always use the tiled DGEMM path                                */
#define SBP_DENSE_NO_SYMMETRY 2 /* do not exploit centro-antisymmetry even if D has it                 */
/* Centro-antisymmetric operators (D[N-1-i][N-1-j] == -D[i][j]: Legendre bases on symmetric nodes, classical FD-SBP)
 * run on the even/odd kernel (half the FP64 work).  The symmetry is accepted when |D + J D J|_max <= 256 eps |D|_max,
 * i.e. also for a matrix that has it only up to the rounding of its construction, such as the barycentric matrix
 * `basis.D` of PolynomialBases.jl that src/polynomialbases_operators.jl:148 multiplies by; the antisymmetric part
 * (D - J D J)/2 is then used, which changes entries by at most 128 eps |D|_max (results move by ~1e-14 relative at
 * N = 64, far inside the 1e-12 contract).  SBP_DENSE_SYMMETRIZE (the former opt-in) is accepted and implied;
 * SBP_DENSE_SYM_BITWISE takes the fast path only when the symmetry holds bitwise (same operator, other summation order). */
#define SBP_DENSE_SYMMETRIZE 4
#define SBP_DENSE_SYM_BITWISE 16
/* Small dense operators whose (4-column x 8-row) fragments are >= 25% structurally zero (banded FSBP / FD-SBP
 * operators stored dense, block-diagonal sub-cell operators) skip those fragments; exact for finite inputs
 * (0*x == 0), differs for Inf/NaN inputs exactly like the sparse path.  This flag multiplies the stored zeros. */
#define SBP_DENSE_NO_SKIP 8

/* ---- library / context ----------------------------------------------------------------- */
int32_t sbp_version(void);
const char* sbp_last_error(void);
int32_t sbp_device_count(int32_t* n);
int32_t sbp_ctx_create(int32_t device, sbp_ctx** ctx);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the ctx's own. */
int32_t sbp_ctx_set_stream(sbp_ctx* ctx, void* cuda_stream);
int32_t sbp_ctx_get_stream(sbp_ctx* ctx, void** cuda_stream);
int32_t sbp_ctx_sync(sbp_ctx* ctx);
int32_t sbp_ctx_destroy(sbp_ctx* ctx);
/* sm count, L2 bytes, max dynamic smem per block, total global memory */
int32_t sbp_ctx_device_info(sbp_ctx* ctx, int32_t* sm_count, int64_t* l2_bytes, int64_t* smem_optin,
                            int64_t* global_bytes);
/* number of kernels this ctx has launched since creation (for harness accounting) */
int64_t sbp_ctx_launch_count(sbp_ctx* ctx);
/* PCI bus id of the ctx's device ("0000:1b:00.0", lower case, NUL-terminated; len >= 16): lets the host side place its pinned
 * buffers and threads on the GPU's NUMA node (/sys/bus/pci/devices/<id>/local_cpulist) before calling the *_host entry points */
int32_t sbp_ctx_pci_bus_id(sbp_ctx* ctx, char* buf, int32_t len);

/* Tuning / debugging options of a context ("bsr_v3", "bsr_pf", "solve_small", "sparse_kernel", ...: the table in
 * csrc/common.cuh, enum Tune).  Every option takes its default from its SBP_* environment variable ONCE, when the context is
 * created; no launch path reads the environment. */
int32_t sbp_ctx_set_option(sbp_ctx* ctx, const char* name, int32_t value);
int32_t sbp_ctx_get_option(sbp_ctx* ctx, const char* name, int32_t* value);

/* ---- memory ---------------------------------------------------------------------------- */
int32_t sbp_malloc(sbp_ctx* ctx, size_t bytes, void** dptr);
int32_t sbp_free(sbp_ctx* ctx, void* dptr);
int32_t sbp_host_alloc(size_t bytes, void** hptr); /* pinned host memory */
int32_t sbp_host_free(void* hptr);
int32_t sbp_memcpy_h2d(sbp_ctx* ctx, void* dptr, const void* h_src, size_t bytes); /* async on ctx stream */
int32_t sbp_memcpy_d2h(sbp_ctx* ctx, void* h_dst, const void* dptr, size_t bytes); /* synchronises */
int32_t sbp_memset(sbp_ctx* ctx, void* dptr, int32_t byte, size_t bytes);

/* ---- operators (uploaded once; construction stays on the CPU) ----------------------------- */

/* Dense N x N operator, `h_D` column-major with leading dimension ldd (Julia Matrix), optional
 * quadrature weights h_w (N, may be NULL -> no integrate/energy support) and a constant `scale`
 * folded into every apply (the `jac` of src/polynomialbases_operators.jl:148).
 * Replaces the storage side of PolynomialBasesDerivativeOperator (src/polynomialbases_operators.jl:7-31),
 * upstream MatrixDerivativeOperator (built at ext/function_space_operators.jl:13-14) and one
 * direction D[dim] of MultidimensionalMatrixDerivativeOperator
 * (ext/multidimensional_function_space_operators.jl:20-22). */
int32_t sbp_op_dense_create(sbp_ctx* ctx, int32_t N, const double* h_D, int32_t ldd, const double* h_w,
                            double scale, int32_t flags, sbp_op** op);

/* CSR operator (0-based, int32).  For operators whose dense storage in the reference is
 * structurally sparse (banded FSBP, ellipsoid-neighbourhood MFSBP: src/utils/sparsity_patterns.jl:48-60). */
int32_t sbp_op_sparse_create(sbp_ctx* ctx, int32_t N, const int32_t* h_rowptr, const int32_t* h_col,
                             const double* h_val, const double* h_w, double scale, sbp_op** op);
/* Same, extracting the exact non-zeros of a dense column-major matrix. */
int32_t sbp_op_sparse_from_dense(sbp_ctx* ctx, int32_t N, const double* h_D, int32_t ldd, const double* h_w,
                                 double scale, sbp_op** op);

/* Table of G dense N x N operators with per-operator weights (h_D: G matrices, each column-major
 * N x N; h_w: G*N).  Serves SubcellOperator families with varying split point: the caller passes
 * D_g = inv(Diagonal(w_L (+) w_R)) * (Q_left + Q_right) (src/subcell_operators.jl:183-186), which the
 * reference rebuilds on every mul! (src/subcell_operators.jl:272). */
int32_t sbp_op_table_create(sbp_ctx* ctx, int32_t N, int32_t G, const double* h_D, const double* h_w,
                            sbp_op** op);

/* 2D linear advection bundle (MultidimensionalLinearAdvectionNonperiodicSemidiscretization,
 * src/conservation_laws/multidimensional_linear_advection.jl:22-41).  `A` = a1*Dx + a2*Dy as a dense or
 * sparse op (kept referenced; destroy it after the bundle), h_b = diagonal of B = inv(P)*(a1*B_x+a2*B_y)
 * (N), h_w = diagonal of P (N), h_mode (N): 0 interior, 1 inflow, 2 outflow -- the state left by the
 * sequential set_bc! loop (:60-71, last writer wins at corners).  Boundary entry lists (length Nb, with
 * duplicated corner nodes) feed energy_rate_boundary (:98-105): h_bnd_node (0-based node of entry i) and
 * h_bnd_tau[i] = weights_boundary[i] * dot(normals[i], a). */
int32_t sbp_op_advection2d_create(sbp_ctx* ctx, const sbp_op* A, const double* h_b, const double* h_w,
                                  const int32_t* h_mode, int32_t Nb, const int32_t* h_bnd_node,
                                  const double* h_bnd_tau, sbp_op** op);

/* 1D variable-coefficient advection bundle (upstream VariableLinearAdvectionNonperiodicSemidiscretization,
 * used at examples/RBF_FSBP_advection.jl:37-38): du = -D*(a.*u) + SATs at first/last node.
 * `D` dense or sparse op with weights; h_a (N) nodal speeds. */
int32_t sbp_op_advection1d_create(sbp_ctx* ctx, const sbp_op* D, const double* h_a, sbp_op** op);

int32_t sbp_op_destroy(sbp_op* op);
/* kind, N, table size G (1 otherwise), nnz (N*N for dense), kernel variant chosen (SBP_KERNEL_*) */
int32_t sbp_op_info(const sbp_op* op, int32_t* kind, int32_t* N, int32_t* G, int64_t* nnz, int32_t* variant);

#define SBP_KERNEL_DMMA_SMEM 1     /* operator fragments resident in shared memory, DMMA tiles         */
#define SBP_KERNEL_DMMA_SMEM_SYM 2 /* same, even/odd split of a centro-antisymmetric operator          */
#define SBP_KERNEL_DGEMM_TILED 3   /* large N: tiled DMMA DGEMM, operator streamed through L2          */
#define SBP_KERNEL_CSR_SMEM 4      /* sparse: instance tile staged in smem, CSR rows per thread        */
#define SBP_KERNEL_TABLE 5
#define SBP_KERNEL_BSR_DMMA 6      /* sparse: 8x4 operator fragments on the FP64 tensor cores, node ring in smem */

/* ---- hot path --------------------------------------------------------------------------- */

/* dU[:,b] = alpha * scale * D * U[:,b] + beta * dU[:,b]   for b in [0,B).
 * Replaces LinearAlgebra.mul!(dest, D, u, alpha, beta) of src/polynomialbases_operators.jl:139-149,
 * src/subcell_operators.jl:270-280 and the upstream MatrixDerivativeOperator / D[dim] methods, applied to
 * every instance of the batch.  N is implied by the operator; U and dU hold B*N doubles; they must not
 * alias.  With beta == 0 dU is write-only (NaNs in dU are not propagated, as in BLAS). */
int32_t sbp_apply(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B, double alpha,
                  double beta);

/* Operator-table apply: instance b uses operator op_index[b] (device int32 array, length B), or, when
 * op_index == NULL, operator (b % G).  Optionally fuses the discrete energy integrate(abs2, u, D_g) =
 * sum_i w_g[i] u_i^2 (src/subcell_operators.jl:92-94) per instance (energy, B doubles or NULL) and its
 * sum over the batch (energy_sum, one device double or NULL; overwritten). Also valid for DENSE ops (G=1). */
int32_t sbp_apply_table(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B,
                        const int32_t* op_index, double alpha, double beta, double* energy,
                        double* energy_sum);

/* out[b] = sum_i w_i * f(U[i,b], V[i,b]) with mode 0: u, 1: u^2 (abs2), 2: u*v, 3: |u| (abs), 16 + k: u^k for an
 * integer 0 <= k < 64 (x -> x^k, evaluated by binary powering like Julia's power_by_squaring); out_sum (device, may be
 * NULL) receives the sum over the batch.  Replaces integrate(func, u, D) (src/polynomialbases_operators.jl:66-68,
 * src/subcell_operators.jl:92-94; the reference accepts any `func`, the device these) and the P-weighted sums of
 * analyze_quantities.  V may be NULL except for mode 2. */
int32_t sbp_integrate(sbp_ctx* ctx, const sbp_op* op, const double* U, const double* V, int64_t B,
                      int32_t mode, double* out, double* out_sum);

/* U[i,b] <- factor * U[i,b] * w_i  (inverse == 0)   or   factor * U[i,b] / w_i  (inverse != 0).
 * Replaces scale_by_mass_matrix! / scale_by_inverse_mass_matrix!
 * (src/polynomialbases_operators.jl:82-114, src/subcell_operators.jl:190-220). */
int32_t sbp_scale_by_mass_matrix(sbp_ctx* ctx, const sbp_op* op, double* U, int64_t B, double factor,
                                 int32_t inverse);

/* Fused 2D advection RHS (functor at src/conservation_laws/multidimensional_linear_advection.jl:73-81 with
 * set_bc! :60-71 folded in):  dU[j,b] = -sum_k A[j,k] U[k,b] + (mode[j]==inflow ? b_j*(U[j,b]-g_j) : 0).
 * g holds the boundary data bc(x_j, t) per NODE (N doubles; only inflow entries are read), shared by the
 * batch when g_stride == 0 or per instance at g + b*g_stride.  quantities (device, 7*B doubles or NULL)
 * receives analyze_quantities (:83-108) per instance in the reference's order
 * [mass, mass_rate, mass_rate_boundary, energy, energy_rate, energy_rate_boundary,
 *  energy_rate_boundary_dissipation]. */
int32_t sbp_rhs_advection2d(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B,
                            const double* g, int64_t g_stride, double* quantities);

/* Fused 1D advection RHS with Godunov SATs (upstream semidiscretization; §3.3 of SURVEY.md):
 * dU = -D*(a.*U); dU[0] += (fl - a0 U0)/w0; dU[N-1] -= (fr - aN UN)/wN with fl = godunov(bc_left, U0, a0),
 * fr = godunov(UN, bc_right, aN).  bc (device): 2 doubles [left,right] shared (bc_stride == 0) or per
 * instance at bc + b*bc_stride.  quantities as above (src/conservation_laws/analysis_callback.jl:136-173). */
int32_t sbp_rhs_advection1d(sbp_ctx* ctx, const sbp_op* op, double* dU, const double* U, int64_t B,
                            const double* bc, int64_t bc_stride, double* quantities);

/* Runge-Kutta stage updates riding on the RHS pass (device-resident time stepping: SSPRK53 of
 * examples/RBF_FSBP_advection.jl:45-51, the low-storage schemes of examples/RBF_MFSBP_advection.jl:22-27):
 *     Unext = ca*U + cb*Uextra + ch*RHS(U),
 * RHS as in sbp_rhs_advection2d / sbp_rhs_advection1d (same g / bc arguments).  Uextra may be NULL (cb ignored) and may
 * alias Unext; Unext must not alias U.  On the block-sparse kernel the update is part of the RHS kernel (U's own rows are
 * already on chip: 16-24 B/DOF per stage instead of 16 + 24..32); other kernel families evaluate the RHS into a library
 * buffer and combine in one pass. */
int32_t sbp_rhs_advection2d_stage(sbp_ctx* ctx, const sbp_op* op, double* Unext, const double* U, int64_t B,
                                  const double* g, int64_t g_stride, double ca, double cb, const double* Uextra,
                                  double ch);
int32_t sbp_rhs_advection1d_stage(sbp_ctx* ctx, const sbp_op* op, double* Unext, const double* U, int64_t B,
                                  const double* bc, int64_t bc_stride, double ca, double cb, const double* Uextra,
                                  double ch);

/* Y = a*X + b*Y + c*Z elementwise over n doubles (Z may be NULL): stage updates of the explicit
 * Runge-Kutta schemes the examples use (SSPRK53, RDPK3SpFSAL49) so that a step never leaves the device. */
int32_t sbp_axpbypcz(sbp_ctx* ctx, double* Y, const double* X, const double* Z, int64_t n, double a,
                     double b, double c);
/* Y = a*X + b*Y + c*Z + d*W in one pass (Z, W may be NULL; Y is not read when b == 0): a whole stage update
 * u3 = a30*u + a32*u2 + b32*h*f(u2) of SSPRK53 (examples/RBF_FSBP_advection.jl:45-51) reads three vectors once. */
int32_t sbp_lincomb(sbp_ctx* ctx, double* Y, const double* X, const double* Z, const double* W, int64_t n, double a,
                    double b, double c, double d);

/* ---- multi-GPU: one process per GPU; the only collective is a sum of a few scalars ------------- */
#define SBP_NCCL_ID_BYTES 128
int32_t sbp_comm_unique_id(char id[SBP_NCCL_ID_BYTES]);       /* rank 0, then broadcast by the host runtime */
int32_t sbp_comm_init(sbp_ctx* ctx, int32_t nranks, int32_t rank, const char id[SBP_NCCL_ID_BYTES]);
int32_t sbp_allreduce_sum(sbp_ctx* ctx, double* dptr, int32_t count); /* ncclAllReduce(double, sum), in place */
int32_t sbp_comm_destroy(sbp_ctx* ctx);

/* ---- device-resident time stepping ------------------------------------------------------------------- */
/* Fixed-step SSPRK53 (`solve(ode, SSPRK53(); dt, adaptive = false)`, examples/RBF_FSBP_advection.jl:45-51) on a batch that
 * never leaves the device: nsteps steps of sizes h_steps[s] (HOST array), five sbp_rhs_advection{2d,1d}_stage launches per
 * step.  `op` is a 2D or 1D advection bundle; U (device, B*N) holds u(t0) on entry and u(t_end) on return.  bc_table
 * (device): boundary data of every stage, row 5*s + k = data of stage k of step s at time t_s + c_k h_s
 * (c = 0, 0.377268915331368, 0.754537830662736, 0.728985661612188, 0.699226135931670), rows bc_row_stride doubles apart,
 * each holding what sbp_rhs_advection2d takes as `g` (N doubles, shared by the batch) resp. sbp_rhs_advection1d as `bc`
 * (2 doubles).  work: 3*B*N doubles of device scratch, or NULL (library-owned, grown on demand). */
int32_t sbp_solve_ssprk53(sbp_ctx* ctx, const sbp_op* op, double* U, int64_t B, const double* bc_table,
                          int64_t bc_row_stride, int64_t nsteps, const double* h_steps, double* work);

/* Low-storage 3S*+ Runge-Kutta schemes with embedded error estimator and FSAL (Ranocha, Dalcin, Parsani, Ketcheson 2021):
 * the family of RDPK3SpFSAL49, the integrator of examples/RBF_MFSBP_advection.jl:21-27.  The scheme is passed as a tableau
 * (the coefficients of RDPK3SpFSAL49 live in OrdinaryDiffEqLowStorageRK; the Julia shim extracts them from the algorithm's
 * cache): one step is  tmp = uprev; u = tmp + beta[0] dt k1; for i = 1..s-1: k = f(u, t + c[i-1] dt); tmp += delta[i-1] u;
 * u = gamma1[i-1] u + gamma2[i-1] tmp + gamma3[i-1] uprev + beta[i] dt k;  error estimate utilde = dt (sum_i bhat[i] k_i +
 * bhat_fsal f(u_new)). */
typedef struct {
    int32_t stages;               /* s                                                                    */
    int32_t order_for_controller; /* k = min(order, embedded order) + 1: exponent divisor of the PID controller */
    const double* gamma1;         /* s - 1 entries each (stages 2..s): gamma1, gamma2, gamma3, delta, c    */
    const double* gamma2;
    const double* gamma3;
    const double* delta;
    const double* c;
    const double* beta;           /* s entries                                                             */
    const double* bhat;           /* s entries (weights of the error estimate)                              */
    double bhat_fsal;
} sbp_tableau_3sp;
/* boundary data at time t into h_row (host; N doubles for a 2D bundle: g per node, 2 doubles [left, right] for a 1D bundle) */
typedef void (*sbp_bc_fn)(double t, double* h_row, void* user);
/* called (stream synchronised, U holds u(t)) whenever the integration lands on a stop time */
typedef void (*sbp_stop_fn)(double t, void* user);
/* Integrate U (device, B*N) from t0 to t1.  adaptive != 0: step size control as OrdinaryDiffEq's PIDController(pid[0..2]) with
 * EEst = rms(utilde / (abstol + max(|uprev|, |u|) reltol)) over the whole batch (one dt for all instances), accept when the
 * limited factor 1 + atan(q - 1) >= 0.81; adaptive == 0: fixed steps dt0.  Steps are clipped to land on tstops (sorted; the
 * tstops of an AnalysisCallback(semi; dt = ...)) and on t1.  h_stats (may be NULL): [accepted, rejected, last dt]. */
int32_t sbp_solve_3sp_fsal(sbp_ctx* ctx, const sbp_op* op, double* U, int64_t B, const sbp_tableau_3sp* tab, double t0,
                           double t1, double dt0, int32_t adaptive, double abstol, double reltol, const double* pid,
                           sbp_bc_fn bc, void* bc_user, const double* tstops, int32_t n_tstops, sbp_stop_fn on_stop,
                           void* stop_user, int64_t max_steps, double* h_stats);
/* rms_i( utilde_i / (abstol + max(|uprev_i|, |u_i|) reltol) ) over n device doubles -> *h_rms (host; synchronises): the error
 * norm of OrdinaryDiffEq's calculate_residuals + ODE_DEFAULT_NORM, deterministic two-pass reduction. */
int32_t sbp_error_norm(sbp_ctx* ctx, const double* utilde, const double* uprev, const double* u, int64_t n, double abstol,
                       double reltol, double* h_rms);

/* ---- operator construction: batched objective evaluation ----------------------------------- */
/* f(x) of the function-space-operator construction for C candidate parameter vectors at once (one CTA per candidate).
 * Reference, one x per call: optimization_function_function_space_operator (ext/function_space_operators_optim.jl:89-105):
 *   f(x) = sum |S(sigma) V - P(rho) V_x + R|^2, x = [sigma (L); rho (N)], P = diag(sig.(rho)) vol / sum(sig.(rho)), sig = logistic;
 * and optimization_function_multidimensional_function_space_operator (ext/multidimensional_function_space_operators.jl:124-161):
 *   f(x) = sum_d ( |S_d V - P V_xd + B_d(phi) V / 2|^2 + |V' B_d(phi) V - M_d|^2 ), x = [sigma_1; ..; sigma_D; rho (N); phi (Nb)].
 * S_d is given by its index map (what set_S! of ext/utils.jl:83-148 writes) as CSR rows per direction: row i of S_d has the entries
 * sign[j] * x[xidx[j]] in column col[j], j in [rowptr[d*(N+1)+i], rowptr[d*(N+1)+i+1]).  V: N x K, V_x: D x N x K, R: D x N x K
 * (constant part of the residual, NULL = 0), M: D x K x K moments (NULL: no moment term), all row-major; boundary map (NULL for
 * the 1D objective): bidx[d*N+k] = index j into phi of the boundary entry that set_B! leaves at node k (-1: none),
 * bn[d*N+k] = normals[j][d].  X: device, C rows of Ltot + N + Nb doubles; out: device, C doubles. */
typedef struct sbp_objective sbp_objective;
int32_t sbp_objective_create(sbp_ctx* ctx, int32_t N, int32_t K, int32_t D, int32_t Nb, int64_t Ltot, const int32_t* h_rowptr,
                             const int32_t* h_col, const int32_t* h_xidx, const double* h_sign, const double* h_V,
                             const double* h_Vx, const double* h_R, const double* h_M, const int32_t* h_bidx,
                             const double* h_bn, double vol, sbp_objective** out);
int32_t sbp_objective_eval(sbp_ctx* ctx, const sbp_objective* obj, const double* X, int64_t C, double* out);
int32_t sbp_objective_destroy(sbp_objective* obj);

/* ---- host-buffer entry points (end-to-end path: pinned staging + chunked H2D/compute/D2H overlap) ------- */
/* h_dU = alpha*scale*D*h_U + beta*h_dU for host arrays; copies are pipelined in chunks of
 * `chunk_instances` (0 = auto) over three streams.  This is what `mul!(dU::Matrix, D, U::Matrix)` on
 * host arrays maps to. */
int32_t sbp_apply_host(sbp_ctx* ctx, const sbp_op* op, double* h_dU, const double* h_U, int64_t B,
                       double alpha, double beta, int64_t chunk_instances);
/* The 2D advection RHS `semi(du, u, p, t)` (src/conservation_laws/multidimensional_linear_advection.jl:73-81) for HOST
 * matrices: h_g holds the boundary data as for sbp_rhs_advection2d but in host memory (N doubles when g_stride == 0, else
 * row b at h_g + b*g_stride); same chunked pipeline as sbp_apply_host. */
int32_t sbp_rhs_advection2d_host(sbp_ctx* ctx, const sbp_op* op, double* h_dU, const double* h_U, int64_t B,
                                 const double* h_g, int64_t g_stride, int64_t chunk_instances);
/* sbp_solve_ssprk53 for a HOST state and HOST boundary table: one upload of u(t0) and the table, nsteps steps on the device,
 * one download of u(t_end) -- the end-to-end form of a whole solve. */
int32_t sbp_solve_ssprk53_host(sbp_ctx* ctx, const sbp_op* op, double* h_U, int64_t B, const double* h_bc_table,
                               int64_t bc_row_stride, int64_t nsteps, const double* h_steps);
/* Pinned-memory cudaMemcpyAsync rates of this box (GB/s): H2D alone, D2H alone, both directions at once (sum) -- the
 * ceiling of the *_host entry points (they move 8 B/DOF each way). */
int32_t sbp_pcie_probe(sbp_ctx* ctx, size_t bytes, double* h2d_gbs, double* d2h_gbs, double* bidir_gbs);

#ifdef __cplusplus
}
#endif
#endif /* SBP_B200_H */
