// objective.cu -- SURVEY 8 (f4): batched evaluation of the operator-CONSTRUCTION objective for many candidate parameter
// vectors x at once (a population of an evolutionary / multi-start optimiser, the trial points of a line search).
//
// Reference (CPU, one x per call, ForwardDiff through it): ext/function_space_operators_optim.jl:89-105
//     f(x) = sum |S(sigma) V - P(rho) V_x + R|^2,   x = [sigma; rho],  P = diag(sig.(rho)) * x_length / sum(sig.(rho)), R = B V / 2
// and ext/multidimensional_function_space_operators.jl:124-161
//     f(x) = sum_d ( |S_d(sigma_d) V - P V_xd + B_d(phi) V / 2|^2 + |V' B_d(phi) V - M_d|^2 ),   x = [sigma_1; ..; sigma_D; rho; phi]
// with sig(x) = 1 / (1 + exp(-x)) (ext/utils.jl:153), S_d skew-symmetric with its free entries taken from sigma_d through an
// index map (set_S!, ext/utils.jl:83-148: full / block-banded / sparsity pattern -- the map is built on the host and uploaded as
// CSR rows {column, index into x, sign}), B_d = diagonal, b_d[k] = sig_b(phi[j]) * normals[j][d] at boundary node k
// (set_B!, ext/utils.jl:179-190; sig_b(x) = x).  The 1D objective is the D = 1 case with a constant R and no moment term.
//
// One CTA per candidate: x, V, the V_xd and the diagonal of P live in shared memory; a thread owns a row of the residual A_d.
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "ptx.cuh"

struct sbp_objective {
    sbp_ctx* ctx = nullptr;
    int N = 0, K = 0, D = 0, Nb = 0;
    long long xlen = 0, Ltot = 0;
    double vol = 1.0;
    int has_R = 0, has_M = 0;
    int *d_rowptr = nullptr, *d_col = nullptr, *d_xidx = nullptr, *d_bidx = nullptr;  // CSR per direction: D * (N + 1) row pointers
    double *d_sign = nullptr, *d_V = nullptr, *d_Vx = nullptr, *d_R = nullptr, *d_M = nullptr, *d_bn = nullptr;
};

namespace sbp {

struct ObjParams {
    int N, K, D, Nb;
    long long xlen, Ltot, C;
    double vol;
    const int* __restrict__ rowptr;
    const int* __restrict__ col;
    const int* __restrict__ xidx;
    const double* __restrict__ sign;
    const double* __restrict__ V;    // N x K row-major
    const double* __restrict__ Vx;   // D x N x K
    const double* __restrict__ R;    // D x N x K or nullptr
    const double* __restrict__ M;    // D x K x K or nullptr
    const int* __restrict__ bidx;    // D x N: index into phi, -1: no boundary term
    const double* __restrict__ bn;   // D x N: normal component
    const double* __restrict__ X;
    double* __restrict__ out;
};

__global__ void __launch_bounds__(128) k_objective(const ObjParams p) {
    extern __shared__ double sm[];
    __shared__ double red[4];
    const int N = p.N, K = p.K;
    double* x = sm;                 // xlen
    double* V = x + p.xlen;         // N * K
    double* pw = V + (size_t)N * K; // N: diagonal of P
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) V[i] = p.V[i];
    for (long long c = blockIdx.x; c < p.C; c += gridDim.x) {
        __syncthreads();
        const double* xc = p.X + c * p.xlen;
        for (long long i = threadIdx.x; i < p.xlen; i += blockDim.x) x[i] = xc[i];
        __syncthreads();
        // P = diag(sig.(rho)) * vol / sum(sig.(rho))  (create_P, ext/utils.jl:164-168)
        const double* rho = x + p.Ltot;
        double s = 0.0;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const double v = 1.0 / (1.0 + exp(-rho[i]));
            pw[i] = v;
            s += v;
        }
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        const double scale = p.vol / (red[0] + red[1] + red[2] + red[3]);
        __syncthreads();
        const double* phi = rho + N;
        double f = 0.0;
        for (int d = 0; d < p.D; d++) {
            const int* rp = p.rowptr + (size_t)d * (N + 1);
            const double* Vx = p.Vx + (size_t)d * N * K;
            // rows of A_d = S_d V - P V_xd + R_d + B_d V / 2
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                const double pi = pw[i] * scale;
                const int bi = p.bidx ? p.bidx[(size_t)d * N + i] : -1;
                const double hb = bi >= 0 ? 0.5 * (phi[bi] * p.bn[(size_t)d * N + i]) : 0.0;
                for (int m = 0; m < K; m++) {
                    double a = 0.0;
                    for (int j = rp[i]; j < rp[i + 1]; j++) a = fma(p.sign[j] * x[p.xidx[j]], V[p.col[j] * K + m], a);
                    a -= pi * Vx[i * K + m];
                    if (p.R) a += p.R[((size_t)d * N + i) * K + m];
                    a = fma(hb, V[i * K + m], a);
                    f = fma(a, a, f);
                }
            }
            // C_d = V' B_d V - M_d  (K x K)
            if (p.M) {
                for (int e = threadIdx.x; e < K * K; e += blockDim.x) {
                    const int m = e / K, m2 = e % K;
                    double cacc = 0.0;
                    for (int i = 0; i < N; i++) {
                        const int bi = p.bidx[(size_t)d * N + i];
                        if (bi >= 0) cacc = fma(V[i * K + m] * (phi[bi] * p.bn[(size_t)d * N + i]), V[i * K + m2], cacc);
                    }
                    cacc -= p.M[((size_t)d * K + m) * K + m2];
                    f = fma(cacc, cacc, f);
                }
            }
        }
        f = warp_sum(f);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = f;
        __syncthreads();
        if (threadIdx.x == 0) p.out[c] = (red[0] + red[1]) + (red[2] + red[3]);
    }
}

template <typename T>
static int32_t up(T** d, const T* h, size_t n) {
    *d = nullptr;
    if (!h || n == 0) return SBP_OK;
    SBP_CUDA(cudaMalloc(d, n * sizeof(T)));
    SBP_CUDA(cudaMemcpy(*d, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return SBP_OK;
}

}  // namespace sbp

using namespace sbp;

extern "C" {

int32_t sbp_objective_destroy(sbp_objective* o) {
    if (!o) return SBP_OK;
    cudaSetDevice(o->ctx->device);
    cudaStreamSynchronize(o->ctx->stream);
    void* ptrs[] = {o->d_rowptr, o->d_col, o->d_xidx, o->d_bidx, o->d_sign, o->d_V, o->d_Vx, o->d_R, o->d_M, o->d_bn};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    delete o;
    return SBP_OK;
}

int32_t sbp_objective_create(sbp_ctx* ctx, int32_t N, int32_t K, int32_t D, int32_t Nb, int64_t Ltot, const int32_t* h_rowptr,
                             const int32_t* h_col, const int32_t* h_xidx, const double* h_sign, const double* h_V,
                             const double* h_Vx, const double* h_R, const double* h_M, const int32_t* h_bidx,
                             const double* h_bn, double vol, sbp_objective** out) {
    SBP_REQUIRE(ctx && h_rowptr && h_V && h_Vx && out, SBP_EINVAL, "null argument");
    SBP_REQUIRE(N >= 1 && K >= 1 && D >= 1 && Nb >= 0 && Ltot >= 0, SBP_EDIM, "bad sizes");
    SBP_REQUIRE((h_M == nullptr && h_bidx == nullptr) || (h_bidx && h_bn), SBP_EINVAL, "boundary map missing");
    SBP_REQUIRE(h_M == nullptr || h_bidx != nullptr, SBP_EINVAL, "the moment term needs the boundary map");
    const int64_t xlen = Ltot + N + Nb;
    for (int d = 0; d < D; d++) {
        const int32_t* rp = h_rowptr + (size_t)d * (N + 1);
        for (int i = 0; i < N; i++) SBP_REQUIRE(rp[i + 1] >= rp[i], SBP_EINVAL, "rowptr of direction %d not monotone", d);
    }
    const int64_t nnz = h_rowptr[(size_t)D * (N + 1) - 1];
    SBP_REQUIRE(nnz == 0 || (h_col && h_xidx && h_sign), SBP_EINVAL, "null entry arrays");
    for (int64_t j = 0; j < nnz; j++)
        SBP_REQUIRE(h_col[j] >= 0 && h_col[j] < N && h_xidx[j] >= 0 && h_xidx[j] < Ltot, SBP_EINVAL, "entry %lld out of range",
                    (long long)j);
    if (h_bidx)
        for (int64_t i = 0; i < (int64_t)D * N; i++)
            SBP_REQUIRE(h_bidx[i] >= -1 && h_bidx[i] < Nb, SBP_EINVAL, "boundary index out of range");
    SBP_REQUIRE((size_t)(xlen + (int64_t)N * K + N) * 8 + 64 <= (size_t)ctx->smem_optin, SBP_EUNSUPPORTED,
                "objective: x (%lld) and V do not fit in shared memory", (long long)xlen);
    SBP_CUDA(cudaSetDevice(ctx->device));
    sbp_objective* o = new sbp_objective();
    o->ctx = ctx, o->N = N, o->K = K, o->D = D, o->Nb = Nb, o->Ltot = Ltot, o->xlen = xlen, o->vol = vol;
    o->has_R = h_R != nullptr, o->has_M = h_M != nullptr;
    int32_t rc = up(&o->d_rowptr, h_rowptr, (size_t)D * (N + 1));
    if (!rc) rc = up(&o->d_col, h_col, (size_t)nnz);
    if (!rc) rc = up(&o->d_xidx, h_xidx, (size_t)nnz);
    if (!rc) rc = up(&o->d_sign, h_sign, (size_t)nnz);
    if (!rc) rc = up(&o->d_V, h_V, (size_t)N * K);
    if (!rc) rc = up(&o->d_Vx, h_Vx, (size_t)D * N * K);
    if (!rc) rc = up(&o->d_R, h_R, (size_t)D * N * K);
    if (!rc) rc = up(&o->d_M, h_M, (size_t)D * K * K);
    if (!rc) rc = up(&o->d_bidx, h_bidx, (size_t)D * N);
    if (!rc) rc = up(&o->d_bn, h_bn, (size_t)D * N);
    if (rc) {
        sbp_objective_destroy(o);
        return rc;
    }
    *out = o;
    return SBP_OK;
}

int32_t sbp_objective_eval(sbp_ctx* ctx, const sbp_objective* o, const double* X, int64_t C, double* out) {
    SBP_REQUIRE(ctx && o, SBP_EINVAL, "null argument");
    SBP_REQUIRE(o->ctx == ctx, SBP_EINVAL, "objective belongs to a different context");
    SBP_REQUIRE(C >= 0, SBP_EDIM, "negative candidate count");
    if (C == 0) return SBP_OK;
    SBP_REQUIRE(X && out, SBP_EINVAL, "null array");
    SBP_CUDA(cudaSetDevice(ctx->device));
    ObjParams p{};
    p.N = o->N, p.K = o->K, p.D = o->D, p.Nb = o->Nb, p.xlen = o->xlen, p.Ltot = o->Ltot, p.C = C, p.vol = o->vol;
    p.rowptr = o->d_rowptr, p.col = o->d_col, p.xidx = o->d_xidx, p.sign = o->d_sign, p.V = o->d_V, p.Vx = o->d_Vx;
    p.R = o->d_R, p.M = o->d_M, p.bidx = o->d_bidx, p.bn = o->d_bn, p.X = X, p.out = out;
    const size_t smem = (size_t)(o->xlen + (int64_t)o->N * o->K + o->N) * sizeof(double);
    if (smem > 48 * 1024) SBP_CUDA(cudaFuncSetAttribute(k_objective, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>(C, (int64_t)ctx->sm_count * 8);
    k_objective<<<grid, 128, smem, ctx->stream>>>(p);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

}  // extern "C"
