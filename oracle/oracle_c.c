/* oracle_c.c -- plain-C restatement of the reference CPU hot path.  TEST INFRASTRUCTURE / CPU BASELINE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product (libsbp_b200.so) never links or calls it.
 *
 * The reference is Julia and cannot be compiled here (no julia binary; see oracle/sbp_oracle.py for the pinning
 * status), so this is a "port": each function restates what one reference call does per instance and loops over
 * the batch the way a Julia user would (`for b in 1:B; mul!(view(dU,:,b), D, view(U,:,b)); end`), optionally
 * spreading the independent instances over OpenMP threads.
 *
 *   oracle_apply_dense      LinearAlgebra.mul!(dest, D::Matrix, u, alpha, beta) -> BLAS dgemv('N') on a column-major
 *                           matrix (src/polynomialbases_operators.jl:148, src/subcell_operators.jl:279, upstream
 *                           MatrixDerivativeOperator); axpy-ordered loop = the dgemv_n algorithm
 *   oracle_apply_dense_blas the same per-instance loop calling OpenBLAS dgemv_ itself (what Julia's mul! dispatches to)
 *   oracle_apply_subcell    src/subcell_operators.jl:270-280 INCLUDING the per-call rebuild
 *                           D = inv(Diagonal(w)) * (Q_left + Q_right) (:183-186) that the reference performs
 *   oracle_rhs_advection2d  src/conservation_laws/multidimensional_linear_advection.jl:60-81 INCLUDING the per-call
 *                           materialisation of -A (:78)
 *   oracle_apply_csr        the same products skipping structural zeros (bit-compatible for finite data)
 *   oracle_energy           integrate(abs2, u, D) = sum_i w_i u_i^2 (src/subcell_operators.jl:92-94)
 */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void gemv_n(int N, double alpha, const double* restrict D /* col-major N x N */, const double* restrict u,
                   double beta, double* restrict y, double* restrict tmp) {
    /* dgemv_n: tmp = D*u accumulated column by column (axpy), then y = alpha*tmp + beta*y */
    for (int i = 0; i < N; i++) tmp[i] = 0.0;
    for (int k = 0; k < N; k++) {
        const double uk = u[k];
        const double* col = D + (size_t)k * N;
        for (int i = 0; i < N; i++) tmp[i] += col[i] * uk;
    }
    if (beta == 0.0)
        for (int i = 0; i < N; i++) y[i] = alpha * tmp[i];
    else
        for (int i = 0; i < N; i++) y[i] = alpha * tmp[i] + beta * y[i];
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_apply_dense(const double* D, int N, const double* U, double* Y, long B, double alpha, double beta,
                        int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        double* tmp = (double*)malloc(sizeof(double) * (size_t)N);
#pragma omp for schedule(static)
        for (long b = 0; b < B; b++) gemv_n(N, alpha, D, U + b * N, beta, Y + b * N, tmp);
        free(tmp);
    }
}

/* The same loop through the BLAS the reference itself calls: `mul!(dest, D::Matrix, u, alpha, beta)` is LinearAlgebra's
 * BLAS.gemv!('N', alpha, D, u, beta, dest) = OpenBLAS dgemv_ (Fortran interface).  `dgemv` is the address of that symbol in
 * the OpenBLAS that numpy / scipy bundle here (taken from scipy.linalg.cython_blas by oracle/c_oracle.py). */
typedef void (*dgemv_fn)(char* trans, int* m, int* n, double* alpha, double* a, int* lda, double* x, int* incx, double* beta,
                         double* y, int* incy);
void oracle_apply_dense_blas(void* dgemv, const double* D, int N, const double* U, double* Y, long B, double alpha,
                             double beta, int nthreads) {
    dgemv_fn f = (dgemv_fn)dgemv;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (long b = 0; b < B; b++) {
        char tr = 'N';
        int n = N, one = 1;
        double a = alpha, be = beta;
        f(&tr, &n, &n, &a, (double*)D, &n, (double*)(U + b * N), &one, &be, Y + b * N, &one);
    }
}

/* op_index == NULL -> operator b % G.  Q_left/Q_right: G matrices col-major N x N; w: G*N (w_L (+) w_R). */
void oracle_apply_subcell(const double* Q_left, const double* Q_right, const double* w, int N, int G,
                          const int* op_index, const double* U, double* Y, double* energy, long B, double alpha,
                          double beta, int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        double* Dm = (double*)malloc(sizeof(double) * (size_t)N * N);
        double* tmp = (double*)malloc(sizeof(double) * (size_t)N);
#pragma omp for schedule(static)
        for (long b = 0; b < B; b++) {
            const int g = op_index ? op_index[b] : (int)(b % G);
            const double* QL = Q_left + (size_t)g * N * N;
            const double* QR = Q_right + (size_t)g * N * N;
            const double* wg = w + (size_t)g * N;
            /* D = Matrix(Dop) rebuilt on every call (:272) */
            for (int k = 0; k < N; k++)
                for (int i = 0; i < N; i++) Dm[(size_t)k * N + i] = (1.0 / wg[i]) * (QL[(size_t)k * N + i] + QR[(size_t)k * N + i]);
            gemv_n(N, alpha, Dm, U + b * N, beta, Y + b * N, tmp);
            if (energy) {
                double e = 0.0;
                for (int i = 0; i < N; i++) e += U[b * N + i] * U[b * N + i] * wg[i];
                energy[b] = e;
            }
        }
        free(Dm);
        free(tmp);
    }
}

void oracle_apply_csr(const int* rowptr, const int* col, const double* val, int N, const double* U, double* Y, long B,
                      double alpha, double beta, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (long b = 0; b < B; b++) {
        const double* u = U + b * N;
        double* y = Y + b * N;
        for (int i = 0; i < N; i++) {
            double s = 0.0;
            for (int j = rowptr[i]; j < rowptr[i + 1]; j++) s += val[j] * u[col[j]];
            y[i] = (beta == 0.0) ? alpha * s : alpha * s + beta * y[i];
        }
    }
}

/* mode: 0 interior, 1 inflow, 2 outflow (state left by set_bc!); g: N doubles shared (g_stride 0) or per instance.
 * faithful != 0: dense A, with -A materialised on every call like the reference; else CSR. */
void oracle_rhs_advection2d(const double* A_dense, const int* rowptr, const int* col, const double* val, int N,
                            const double* bdiag, const int* mode, const double* g, long g_stride, const double* U,
                            double* dU, long B, int faithful, int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        double* negA = faithful ? (double*)malloc(sizeof(double) * (size_t)N * N) : NULL;
        double* tmp = (double*)malloc(sizeof(double) * (size_t)N);
        double* tmp1 = (double*)malloc(sizeof(double) * (size_t)N);
#pragma omp for schedule(static)
        for (long b = 0; b < B; b++) {
            const double* u = U + b * N;
            double* du = dU + b * N;
            const double* gb = g + b * g_stride;
            for (int i = 0; i < N; i++) tmp1[i] = mode[i] == 1 ? gb[i] : (mode[i] == 2 ? u[i] : 0.0); /* set_bc! */
            if (faithful) {
                for (size_t e = 0; e < (size_t)N * N; e++) negA[e] = -A_dense[e]; /* (-A) allocated per call (:78) */
                gemv_n(N, 1.0, negA, u, 0.0, du, tmp);
            } else {
                for (int i = 0; i < N; i++) {
                    double s = 0.0;
                    for (int j = rowptr[i]; j < rowptr[i + 1]; j++) s += val[j] * u[col[j]];
                    du[i] = -s;
                }
            }
            for (int i = 0; i < N; i++) du[i] = du[i] + bdiag[i] * (u[i] - tmp1[i]);
        }
        if (negA) free(negA);
        free(tmp);
        free(tmp1);
    }
}

void oracle_energy(const double* w, int N, int G, const double* U, double* energy, long B, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (long b = 0; b < B; b++) {
        const double* wg = w + (size_t)(G > 1 ? b % G : 0) * N;
        double e = 0.0;
        for (int i = 0; i < N; i++) e += U[b * N + i] * U[b * N + i] * wg[i];
        energy[b] = e;
    }
}

/* Whole fixed-step SSPRK53 solve of the upstream 1D semidiscretization for ONE instance, as examples/RBF_FSBP_advection.jl
 * runs it on the CPU (BASELINE config 1): per stage du = -D (a .* u) through the dense dgemv the reference's operator type
 * performs (MatrixDerivativeOperator mul! on the dense-stored D), Godunov SATs at both ends, then the Shu-Osher combination.
 * bc: (5 * nsteps) rows [left, right] at the stage times; u (N) in/out; work: 5 * N doubles. */
static void rhs1d(int N, const double* restrict D, const double* restrict a, double w0, double wN, const double* restrict u,
                  const double* bc, double* restrict du, double* restrict au, double* restrict tmp) {
    for (int i = 0; i < N; i++) au[i] = a[i] * u[i];
    gemv_n(N, -1.0, D, au, 0.0, du, tmp);
    const double fl = a[0] > 0.0 ? a[0] * bc[0] : a[0] * u[0];
    const double fr = a[N - 1] > 0.0 ? a[N - 1] * u[N - 1] : a[N - 1] * bc[1];
    du[0] += (fl - a[0] * u[0]) / w0;
    du[N - 1] -= (fr - a[N - 1] * u[N - 1]) / wN;
}
void oracle_solve_ssprk53_1d(const double* D, int N, const double* a, double w0, double wN, double* u, long nsteps,
                             const double* hsteps, const double* bc, double* work) {
    const double a30 = 0.355909775063327, a32 = 0.644090224936674, a40 = 0.367933791638137, a43 = 0.632066208361863,
                 a52 = 0.237593836598569, a54 = 0.762406163401431, b10 = 0.377268915331368, b21 = 0.377268915331368,
                 b32 = 0.242995220537396, b43 = 0.238458932846290, b54 = 0.287632146308408;
    double *u1 = work, *u2 = work + N, *u3 = work + 2 * N, *du = work + 3 * N, *au = work + 4 * N;
    double* tmp = (double*)malloc(sizeof(double) * (size_t)N);
    for (long s = 0; s < nsteps; s++) {
        const double h = hsteps[s];
        const double* b = bc + 10 * s;
        rhs1d(N, D, a, w0, wN, u, b, du, au, tmp);
        for (int i = 0; i < N; i++) u1[i] = u[i] + b10 * h * du[i];
        rhs1d(N, D, a, w0, wN, u1, b + 2, du, au, tmp);
        for (int i = 0; i < N; i++) u2[i] = u1[i] + b21 * h * du[i];
        rhs1d(N, D, a, w0, wN, u2, b + 4, du, au, tmp);
        for (int i = 0; i < N; i++) u3[i] = a30 * u[i] + a32 * u2[i] + b32 * h * du[i];
        rhs1d(N, D, a, w0, wN, u3, b + 6, du, au, tmp);
        for (int i = 0; i < N; i++) u1[i] = a40 * u[i] + a43 * u3[i] + b43 * h * du[i];
        rhs1d(N, D, a, w0, wN, u1, b + 8, du, au, tmp);
        for (int i = 0; i < N; i++) u[i] = a52 * u2[i] + a54 * u1[i] + b54 * h * du[i];
    }
    free(tmp);
}
