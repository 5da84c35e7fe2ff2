// Residual-based a-posteriori error estimator of PoissonProblemPrimal, evaluated over all (cell, multi-index)
// and (face, multi-index) pairs on the device (estimate(::Type{PoissonProblemPrimal}, ...), src/estimate.jl:260-418).
//
//   volume part (:308-366)  eta4cell[T,j] = |T|^{3 or 1} sum_q w_q ( f(x_q) [j==1] + [order>1] ( a_0 Lap u_j [j<=N]
//                                            + sum_m a_m(x_q) (g+ Lap u_{j+e_m} + g- Lap u_{j-e_m}) ) )^2
//   jump part   (:371-415)  J[F,j] = int_F | Pi_F sum_{m=0..M_ext} a_m [[grad w_{j,m}]] |^2  (* |F| for j<=N, / |F| else),
//                           w_{j,0} = u_j, w_{j,m} = g+ u_{j+e_m} + g- u_{j-e_m} (active neighbours only),
//                           0 on boundary faces; added to the three faces of every cell and to eta4modes.
// Pi_F = L2 projection onto P_{order-1}(F) with the caller's 1-D rule (assumed FaceInterpolator semantics,
// SURVEY.md B.4).  Neighbour tables / coupling weights of the EXTENDED set come from index.cpp (same code that
// reproduces get_neighbours and G bit-exactly).
//
// Kernels: one CTA per cell (resp. interior face); phase A computes Lap u_k (resp. the gradient jump of u_k at the
// face end points) for all active modes k into shared memory with coalesced reads of the mode-fastest solution
// block; phase B lets every thread own one extended mode j and contract its coupling list against that table -
// the same sum_m diag(a_m) (x) G_m contraction as the operator.  All sums are in a fixed order (deterministic).
#include <algorithm>
#include <cmath>

#include "coeff.cuh"
#include "common.h"

namespace asgfem {

namespace {

constexpr int MAXQ = 64, MAXQF = 16;
__constant__ double e_xref[2 * MAXQ];
__constant__ double e_w[MAXQ];
__constant__ double e_sf[MAXQF];
__constant__ double e_wf[MAXQF];

struct EstArgs {
    int64_t ld, ldE, ncells, nfaces;
    int N, N_ext, M_ext, order, nd, nq, nqf;
    const double* u;
    const int32_t* pos;  // device column of active mode k (private column order of the vectors)
    const double* coords;
    const int32_t *cellnodes, *celldofs;
    const int32_t *cptr, *cm, *ck;
    const double* cg;
    const double* fq;  // nq x ncells or null
    double mean;
    const double* decay;
    const int32_t *b1, *b2;
    const int32_t *face_nodes, *face_cells, *face_loc, *cell_faces;
    double* E;   // ncells x ldE
    double* JF;  // nfaces x ldE
};

__global__ void __launch_bounds__(256) k_est_volume(EstArgs a) {
    extern __shared__ double sm[];
    double* lap = sm;                    // [N]
    double* am = sm + a.N;               // [nq][M_ext+1]
    __shared__ double s_lapphi[6];
    __shared__ double s_vol;
    for (int64_t cell = blockIdx.x; cell < a.ncells; cell += gridDim.x) {
        __syncthreads();
        const int32_t* cn = a.cellnodes + 3 * cell;
        if (threadIdx.x == 0) {
            double gl[3][2];
            double det = lambda_gradients(a.coords, cn, gl);
            s_vol = 0.5 * fabs(det);
            for (int d = 0; d < 6; ++d) s_lapphi[d] = 0.0;
            if (a.order == 2) {
                for (int i = 0; i < 3; ++i) s_lapphi[i] = 4.0 * (gl[i][0] * gl[i][0] + gl[i][1] * gl[i][1]);
                for (int f = 0; f < 3; ++f) {
                    int i = f, j = (f + 1) % 3;
                    s_lapphi[3 + f] = 8.0 * (gl[i][0] * gl[j][0] + gl[i][1] * gl[j][1]);
                }
            }
        }
        // a_m at the quadrature points of this cell
        {
            double x1 = a.coords[2 * cn[0]], y1 = a.coords[2 * cn[0] + 1];
            double x2 = a.coords[2 * cn[1]], y2 = a.coords[2 * cn[1] + 1];
            double x3 = a.coords[2 * cn[2]], y3 = a.coords[2 * cn[2] + 1];
            for (int t = threadIdx.x; t < a.nq * (a.M_ext + 1); t += blockDim.x) {
                int q = t / (a.M_ext + 1), m = t - q * (a.M_ext + 1);
                double xr = e_xref[2 * q], yr = e_xref[2 * q + 1];
                double px = x1 + xr * (x2 - x1) + yr * (x3 - x1);
                double py = y1 + xr * (y2 - y1) + yr * (y3 - y1);
                am[t] = eval_am(m, px, py, a.mean, a.decay, a.b1, a.b2);
            }
        }
        __syncthreads();
        if (a.order > 1) {
            const int32_t* cd = a.celldofs + (int64_t)a.nd * cell;
            for (int k = threadIdx.x; k < a.N; k += blockDim.x) {
                double s = 0.0;
                for (int d = 0; d < a.nd; ++d) s += a.u[(int64_t)cd[d] * a.ld + a.pos[k]] * s_lapphi[d];
                lap[k] = s;
            }
        }
        __syncthreads();
        const double vol = s_vol;
        for (int j = threadIdx.x; j < a.N_ext; j += blockDim.x) {
            double val = 0.0;
            for (int q = 0; q < a.nq; ++q) {
                double r = 0.0;
                if (j == 0) r = a.fq ? a.fq[(int64_t)cell * a.nq + q] : 1.0;
                if (a.order > 1) {
                    const double* amq = am + q * (a.M_ext + 1);
                    if (j < a.N) r += lap[j] * amq[0];
                    for (int e = a.cptr[j]; e < a.cptr[j + 1]; ++e) r += (lap[a.ck[e]] * a.cg[e]) * amq[a.cm[e]];
                }
                val += r * r * e_w[q];
            }
            val *= (j < a.N) ? vol * vol * vol : vol;
            a.E[cell * a.ldE + j] = val;
        }
    }
}

// gradient of basis function d of the cell at barycentric point lam (P1: constant)
__device__ __forceinline__ void grad_phi(int order, int d, const double* lam, const double gl[3][2], double* g2) {
    if (order == 1) {
        g2[0] = gl[d][0];
        g2[1] = gl[d][1];
    } else {
        double dl[3];
        p2_dphi(lam, d, dl);
        g2[0] = dl[0] * gl[0][0] + dl[1] * gl[1][0] + dl[2] * gl[2][0];
        g2[1] = dl[0] * gl[0][1] + dl[1] * gl[1][1] + dl[2] * gl[2][1];
    }
}

__global__ void __launch_bounds__(256) k_est_jumps(EstArgs a) {
    extern __shared__ double sm[];
    // jump of grad u_k at the two face end points: jmp[(k*2 + pt)*2 + comp]
    double* jmp = sm;                       // [N][2][2]
    double* am = sm + (size_t)a.N * 4;      // [nqf][M_ext+1]
    __shared__ double s_g[2][2][6][2];      // side, end point, dof, component
    __shared__ int32_t s_dof[2][6];
    __shared__ double s_len;
    const int npt = a.order == 1 ? 1 : 2;
    for (int64_t face = blockIdx.x; face < a.nfaces; face += gridDim.x) {
        const int32_t c0 = a.face_cells[2 * face], c1 = a.face_cells[2 * face + 1];
        if (c1 < 0) {  // boundary face: jumps4face[bfaces] = 0 (:401)
            for (int j = threadIdx.x; j < a.N_ext; j += blockDim.x) a.JF[face * a.ldE + j] = 0.0;
            continue;
        }
        __syncthreads();
        const int32_t na = a.face_nodes[2 * face], nb = a.face_nodes[2 * face + 1];
        const double ax = a.coords[2 * na], ay = a.coords[2 * na + 1], bx = a.coords[2 * nb], by = a.coords[2 * nb + 1];
        if (threadIdx.x < 2) {
            int side = threadIdx.x;
            int32_t cell = side == 0 ? c0 : c1;
            const int32_t* cn = a.cellnodes + 3 * (int64_t)cell;
            double gl[3][2];
            lambda_gradients(a.coords, cn, gl);
            for (int d = 0; d < a.nd; ++d) s_dof[side][d] = a.celldofs[(int64_t)a.nd * cell + d];
            for (int pt = 0; pt < 2; ++pt) {
                int32_t node = pt == 0 ? na : nb;
                double lam[3] = {cn[0] == node ? 1.0 : 0.0, cn[1] == node ? 1.0 : 0.0, cn[2] == node ? 1.0 : 0.0};
                for (int d = 0; d < a.nd; ++d) grad_phi(a.order, d, lam, gl, s_g[side][pt][d]);
            }
            if (side == 0) s_len = sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
        }
        for (int t = threadIdx.x; t < a.nqf * (a.M_ext + 1); t += blockDim.x) {
            int q = t / (a.M_ext + 1), m = t - q * (a.M_ext + 1);
            double s = e_sf[q];
            am[t] = eval_am(m, ax + s * (bx - ax), ay + s * (by - ay), a.mean, a.decay, a.b1, a.b2);
        }
        __syncthreads();
        for (int k = threadIdx.x; k < a.N; k += blockDim.x) {
            const int kc = a.pos[k];
            for (int pt = 0; pt < npt; ++pt) {
                double gx = 0.0, gy = 0.0;
                for (int d = 0; d < a.nd; ++d) {
                    double u0 = a.u[(int64_t)s_dof[0][d] * a.ld + kc];
                    gx += u0 * s_g[0][pt][d][0];
                    gy += u0 * s_g[0][pt][d][1];
                }
                for (int d = 0; d < a.nd; ++d) {
                    double u1 = a.u[(int64_t)s_dof[1][d] * a.ld + kc];
                    gx -= u1 * s_g[1][pt][d][0];
                    gy -= u1 * s_g[1][pt][d][1];
                }
                jmp[(k * 2 + pt) * 2 + 0] = gx;
                jmp[(k * 2 + pt) * 2 + 1] = gy;
            }
            if (npt == 1) {
                jmp[(k * 2 + 1) * 2 + 0] = jmp[(k * 2) * 2 + 0];
                jmp[(k * 2 + 1) * 2 + 1] = jmp[(k * 2) * 2 + 1];
            }
        }
        __syncthreads();
        const double len = s_len;
        for (int j = threadIdx.x; j < a.N_ext; j += blockDim.x) {
            // projection coefficients onto the orthonormal Legendre basis of P_{order-1}([0,1])
            double c0x = 0.0, c0y = 0.0, c1x = 0.0, c1y = 0.0;
            for (int q = 0; q < a.nqf; ++q) {
                const double s = e_sf[q];
                const double* amq = am + q * (a.M_ext + 1);
                double gx = 0.0, gy = 0.0;
                if (j < a.N) {
                    const double* jp = jmp + (size_t)j * 4;
                    gx += amq[0] * ((1.0 - s) * jp[0] + s * jp[2]);
                    gy += amq[0] * ((1.0 - s) * jp[1] + s * jp[3]);
                }
                for (int e = a.cptr[j]; e < a.cptr[j + 1]; ++e) {
                    const double* jp = jmp + (size_t)a.ck[e] * 4;
                    const double c = amq[a.cm[e]] * a.cg[e];
                    gx += c * ((1.0 - s) * jp[0] + s * jp[2]);
                    gy += c * ((1.0 - s) * jp[1] + s * jp[3]);
                }
                const double wq = e_wf[q];
                c0x += wq * gx;
                c0y += wq * gy;
                if (a.order > 1) {
                    const double L1 = 1.7320508075688772 * (2.0 * s - 1.0);
                    c1x += wq * gx * L1;
                    c1y += wq * gy * L1;
                }
            }
            double val = len * (c0x * c0x + c0y * c0y + c1x * c1x + c1y * c1y);
            val = (j < a.N) ? val * len : val / len;
            a.JF[face * a.ldE + j] = val;
        }
    }
}

// ---- log-transformed primal problem (estimate(::Type{LogTransformedPoissonProblemPrimal}, ...), src/estimate.jl:70-257) ----
//   volume part (:134-217)  eta4cell[T,j] = |T|^{2 or 1} sum_q w_q ( lambda_j(x_q) f(x_q) + [j<=N, order>1] Lap u_j
//                                            + sum_m grad a_m(x_q) . grad w_{j,m}(x_q) )^2,  w_{j,m} = g+ u_{j+e_m} + g- u_{j-e_m}
//   data part   (:156-175)  zeta1 = sum_T |T| sum_q w_q f^2 exp(2 sum_{m<=maxm} a_m^2),  zeta2 = sum_T |T| sum_q w_q f^2 sum_j lambda_j^2
//   jump part   (:232-244)  J[F,j] = |F| int_F |[[grad u_j]]|^2 for the active modes, interior faces
// lambda_j = PCE coefficient of exp(-a) (expa_PCE_mop, factor -1).  The reference interpolates lambda_j into H1Pk{quadorder} and
// evaluates the interpolant at the quadrature points (the direct evaluation is the commented-out "most expensive line" :184);
// here lam_qp (the caller's interpolated values, [cell][q][j]) is used when given, else lambda_j is evaluated directly.
struct EstLogArgs {
    const int32_t* mi;   // N_ext x M_ext
    const double* den;   // sqrt(prod mu_d!) * (-1)^|mu|
    const double* lam_qp;
    int ntrunc;
    double* zeta1;  // [ncells]
    double* Z2;     // ncells x ldE: |T| sum_q w_q f^2 lambda_j^2
};

__global__ void __launch_bounds__(256) k_estlog_volume(EstArgs a, EstLogArgs L) {
    extern __shared__ double sm[];
    double* gu = sm;                       // [N][2]: grad u_k at the current quadrature point
    double* lap = gu + 2 * (size_t)a.N;    // [N]
    double* am = lap + a.N;                // [max(ntrunc, M_ext) + 1]
    const int nam = max(L.ntrunc, a.M_ext);
    double* gam = am + nam + 1;            // [M_ext + 1][2]
    __shared__ double s_gl[3][2], s_lapphi[6], s_vol, s_pref, s_S;
    for (int64_t cell = blockIdx.x; cell < a.ncells; cell += gridDim.x) {
        __syncthreads();
        const int32_t* cn = a.cellnodes + 3 * cell;
        const int32_t* cd = a.celldofs + (int64_t)a.nd * cell;
        const double x1 = a.coords[2 * cn[0]], y1 = a.coords[2 * cn[0] + 1], x2 = a.coords[2 * cn[1]], y2 = a.coords[2 * cn[1] + 1],
                     x3 = a.coords[2 * cn[2]], y3 = a.coords[2 * cn[2] + 1];
        if (threadIdx.x == 0) {
            double gl[3][2];
            const double det = lambda_gradients(a.coords, cn, gl);
            s_vol = 0.5 * fabs(det);
            for (int i = 0; i < 3; ++i) s_gl[i][0] = gl[i][0], s_gl[i][1] = gl[i][1];
            for (int d = 0; d < 6; ++d) s_lapphi[d] = 0.0;
            if (a.order == 2) {
                for (int i = 0; i < 3; ++i) s_lapphi[i] = 4.0 * (gl[i][0] * gl[i][0] + gl[i][1] * gl[i][1]);
                for (int f = 0; f < 3; ++f) {
                    const int i = f, j = (f + 1) % 3;
                    s_lapphi[3 + f] = 8.0 * (gl[i][0] * gl[j][0] + gl[i][1] * gl[j][1]);
                }
            }
        }
        __syncthreads();
        if (a.order > 1)
            for (int k = threadIdx.x; k < a.N; k += blockDim.x) {
                double s = 0.0;
                for (int d = 0; d < a.nd; ++d) s += a.u[(int64_t)cd[d] * a.ld + a.pos[k]] * s_lapphi[d];
                lap[k] = s;
            }
        double acc[8], z2[8];  // extended modes of this thread: j = threadIdx.x + 256 r (N_ext <= 2048)
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = z2[r] = 0.0;
        double zeta1 = 0.0;
        for (int q = 0; q < a.nq; ++q) {
            const double xr = e_xref[2 * q], yr = e_xref[2 * q + 1];
            const double px = x1 + xr * (x2 - x1) + yr * (x3 - x1), py = y1 + xr * (y2 - y1) + yr * (y3 - y1);
            __syncthreads();  // tables of the previous point are no longer read
            for (int m = threadIdx.x + 1; m <= nam; m += blockDim.x) am[m] = eval_am(m, px, py, a.mean, a.decay, a.b1, a.b2);
            for (int m = threadIdx.x + 1; m <= a.M_ext; m += blockDim.x) eval_gradam(m, px, py, a.decay, a.b1, a.b2, gam[2 * m], gam[2 * m + 1]);
            if (a.order > 1 || q == 0) {
                const double lam[3] = {1.0 - xr - yr, xr, yr};
                double gl[3][2];
                for (int i = 0; i < 3; ++i) gl[i][0] = s_gl[i][0], gl[i][1] = s_gl[i][1];
                for (int k = threadIdx.x; k < a.N; k += blockDim.x) {
                    double gx = 0.0, gy = 0.0;
                    for (int d = 0; d < a.nd; ++d) {
                        double g2[2];
                        grad_phi(a.order, d, lam, gl, g2);
                        const double uv = a.u[(int64_t)cd[d] * a.ld + a.pos[k]];
                        gx += uv * g2[0];
                        gy += uv * g2[1];
                    }
                    gu[2 * k] = gx, gu[2 * k + 1] = gy;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                double S = 0.0;
                for (int m = 1; m <= L.ntrunc; ++m) S += am[m] * am[m];
                s_S = S;
                s_pref = exp(S / 2) * exp(-a.mean);
            }
            __syncthreads();
            const double fval = a.fq[(int64_t)cell * a.nq + q], wq = e_w[q], pref = s_pref;
            if (threadIdx.x == 0) zeta1 += fval * fval * exp(2.0 * s_S) * wq * s_vol;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int j = threadIdx.x + 256 * r;
                if (j < a.N_ext) {
                    double lamj;
                    if (L.lam_qp) {
                        lamj = L.lam_qp[((int64_t)cell * a.nq + q) * a.N_ext + j];
                    } else {
                        double amu = 1.0;
                        for (int d = 0; d < a.M_ext; ++d) {
                            const int e = L.mi[(int64_t)j * a.M_ext + d];
                            for (int t = 0; t < e; ++t) amu *= am[d + 1];
                        }
                        lamj = amu / L.den[j] * pref;
                    }
                    double val = lamj * fval;
                    if (a.order > 1 && j < a.N) val += lap[j];
                    double sig = 0.0;
                    for (int e = a.cptr[j]; e < a.cptr[j + 1]; ++e)
                        sig += a.cg[e] * (gam[2 * a.cm[e]] * gu[2 * a.ck[e]] + gam[2 * a.cm[e] + 1] * gu[2 * a.ck[e] + 1]);
                    acc[r] += (val + sig) * (val + sig) * wq;
                    z2[r] += lamj * lamj * fval * fval * wq;
                }
            }
        }
        const double vol = s_vol;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int j = threadIdx.x + 256 * r;
            if (j < a.N_ext) {
                a.E[cell * a.ldE + j] = acc[r] * (j < a.N ? vol * vol : vol);
                L.Z2[cell * a.ldE + j] = z2[r] * vol;
            }
        }
        if (threadIdx.x == 0) L.zeta1[cell] = zeta1;
    }
}

// J[F, k] = |F| int_F |[[grad u_k]]|^2 ds for the active modes on interior faces (jump(grad(1)) of ItemIntegratorDG, :234-241)
__global__ void __launch_bounds__(256) k_estlog_jumps(EstArgs a) {
    __shared__ double s_g[2][2][6][2];
    __shared__ int32_t s_dof[2][6];
    __shared__ double s_len;
    const int npt = a.order == 1 ? 1 : 2;
    for (int64_t face = blockIdx.x; face < a.nfaces; face += gridDim.x) {
        const int32_t c0 = a.face_cells[2 * face], c1 = a.face_cells[2 * face + 1];
        if (c1 < 0) {
            for (int j = threadIdx.x; j < a.N_ext; j += blockDim.x) a.JF[face * a.ldE + j] = 0.0;
            continue;
        }
        __syncthreads();
        const int32_t na = a.face_nodes[2 * face], nb = a.face_nodes[2 * face + 1];
        const double ax = a.coords[2 * na], ay = a.coords[2 * na + 1], bx = a.coords[2 * nb], by = a.coords[2 * nb + 1];
        if (threadIdx.x < 2) {
            const int side = threadIdx.x;
            const int32_t cell = side == 0 ? c0 : c1;
            const int32_t* cn = a.cellnodes + 3 * (int64_t)cell;
            double gl[3][2];
            lambda_gradients(a.coords, cn, gl);
            for (int d = 0; d < a.nd; ++d) s_dof[side][d] = a.celldofs[(int64_t)a.nd * cell + d];
            for (int pt = 0; pt < 2; ++pt) {
                const int32_t node = pt == 0 ? na : nb;
                double lam[3] = {cn[0] == node ? 1.0 : 0.0, cn[1] == node ? 1.0 : 0.0, cn[2] == node ? 1.0 : 0.0};
                for (int d = 0; d < a.nd; ++d) grad_phi(a.order, d, lam, gl, s_g[side][pt][d]);
            }
            if (side == 0) s_len = sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
        }
        __syncthreads();
        const double len = s_len;
        for (int j = threadIdx.x; j < a.N_ext; j += blockDim.x) {
            double val = 0.0;
            if (j < a.N) {
                const int kc = a.pos[j];
                double J[2][2];
                for (int pt = 0; pt < npt; ++pt) {
                    double gx = 0.0, gy = 0.0;
                    for (int d = 0; d < a.nd; ++d) {
                        const double u0 = a.u[(int64_t)s_dof[0][d] * a.ld + kc], u1 = a.u[(int64_t)s_dof[1][d] * a.ld + kc];
                        gx += u0 * s_g[0][pt][d][0] - u1 * s_g[1][pt][d][0];
                        gy += u0 * s_g[0][pt][d][1] - u1 * s_g[1][pt][d][1];
                    }
                    J[pt][0] = gx, J[pt][1] = gy;
                }
                if (npt == 1) J[1][0] = J[0][0], J[1][1] = J[0][1];
                for (int q = 0; q < a.nqf; ++q) {
                    const double s = e_sf[q];
                    const double gx = (1.0 - s) * J[0][0] + s * J[1][0], gy = (1.0 - s) * J[0][1] + s * J[1][1];
                    val += e_wf[q] * (gx * gx + gy * gy);
                }
                val *= len * len;  // integral over the face (length) and jumps4face .*= FaceVolumes
            }
            a.JF[face * a.ldE + j] = val;
        }
    }
}

// partial column sums over row chunks: part[chunk][j] = sum_{r in chunk} A[r][j]   (fixed order)
// wrow (may be null): weight of a row - the share of a cell / face this rank owns in a row-sharded run
__global__ void k_colsum_partial(const double* __restrict__ A, int64_t nrows, int64_t ld, int ncols, int chunk,
                                 double* __restrict__ part, const double* __restrict__ wrow) {
    int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = min(r0 + chunk, nrows);
    double s = 0.0;
    if (wrow) {
        for (int64_t r = r0; r < r1; ++r) s += wrow[r] * A[r * ld + j];
    } else {
        for (int64_t r = r0; r < r1; ++r) s += A[r * ld + j];
    }
    part[(int64_t)blockIdx.x * ncols + j] = s;
}

// cellsum[c] = sum over the selected columns of E[c, .]   (marking indicator sum(eta4cell[:, actives], dims = 2),
// scripts/poisson.jl:402), columns in the given order
__global__ void k_rowsum_selected(const double* __restrict__ E, int64_t ncells, int64_t ldE, const int32_t* __restrict__ sel, int nsel,
                                  double* __restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    double s = 0.0;
    for (int k = 0; k < nsel; ++k) s += E[c * ldE + sel[k]];
    out[c] = s;
}

__global__ void k_colsum_final(const double* __restrict__ part, int nchunks, int ncols, double* __restrict__ out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += part[(int64_t)c * ncols + j];
    out[j] = s;
}

__global__ void k_add_face_jumps(double* __restrict__ E, const double* __restrict__ JF, const int32_t* __restrict__ cell_faces,
                                 int64_t ncells, int64_t ldE, int ncols) {
    for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
        const int32_t f0 = cell_faces[3 * cell], f1 = cell_faces[3 * cell + 1], f2 = cell_faces[3 * cell + 2];
        for (int j = threadIdx.x; j < ncols; j += blockDim.x) {
            double v = E[cell * ldE + j];
            v += JF[(int64_t)f0 * ldE + j];
            v += JF[(int64_t)f1 * ldE + j];
            v += JF[(int64_t)f2 * ldE + j];
            E[cell * ldE + j] = v;
        }
    }
}

// E[rows x ld] (row-major) -> stage[(j - j0) * nrows + r] for a chunk of columns (column-major host layout)
__global__ void k_cols_to_stage(const double* __restrict__ d, double* __restrict__ stage, int64_t nrows, int64_t ld,
                                int64_t j0, int jc) {
    __shared__ double tile[32][33];
    int64_t r0 = (int64_t)blockIdx.x * 32;
    int k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t i = r0 + r;
        int k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < nrows && k < jc) ? d[i * ld + j0 + k] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int k = k0 + r;
        int64_t i = r0 + threadIdx.x;
        if (k < jc && i < nrows) stage[(int64_t)k * nrows + i] = tile[threadIdx.x][r];
    }
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
};

}  // namespace

int estimate_poisson_primal(asgfem_ctx* ctx, const double* u, int64_t N_ext, int64_t M_ext, const int64_t* mi_ext,
                            int32_t nq, const double* xref, const double* w, const double* f_at_qp, int32_t nqf,
                            const double* sf, const double* wf, double* eta4cell, double* eta4modes, int64_t nsel,
                            const int64_t* sel, double* cellsum, int kind, const double* lam_at_qp, int32_t ntrunc, double* zeta3) {
    const int64_t N = ctx->N, Mact = ctx->mis.M, ncells = ctx->ncells;
    // the active modes must be the first N extended modes (padded with zeros) - estimate.jl relies on this
    for (int64_t j = 0; j < N; ++j)
        for (int64_t m = 0; m < M_ext; ++m) {
            int64_t want = m < Mact ? ctx->mis.mi[j * Mact + m] : 0;
            ASG_CHECK(ctx, mi_ext[j * M_ext + m] == want, ASGFEM_EINVAL,
                      "estimate: the first N extended multi-indices must be the active ones");
        }
    // ---- couplings of the extended set restricted to active neighbours (estimate.jl:342-345, 388-395) ----
    MultiIndexSet ext;
    ext.N = N_ext;
    ext.M = M_ext;
    ext.mi.assign(mi_ext, mi_ext + N_ext * M_ext);
    ext.build_neighbours();
    std::vector<double> gp, gm;
    coupling_weights(ctx->family, ext.maxdeg() + 1, gp, gm);
    std::vector<int32_t> cptr((size_t)N_ext + 1, 0), cm, ck;
    std::vector<double> cg;
    for (int64_t j = 0; j < N_ext; ++j) {
        for (int64_t m = 0; m < M_ext; ++m) {
            int64_t deg = ext.mi[j * M_ext + m];
            int64_t p = ext.plus[m + M_ext * j], q = ext.minus[m + M_ext * j];
            if (p > 0 && p <= N) {
                cm.push_back((int32_t)(m + 1));
                ck.push_back((int32_t)(p - 1));
                cg.push_back(gp[deg]);
            }
            if (q > 0 && q <= N) {
                cm.push_back((int32_t)(m + 1));
                ck.push_back((int32_t)(q - 1));
                cg.push_back(gm[deg]);
            }
        }
        cptr[j + 1] = (int32_t)cm.size();
    }
    // ---- faces (own enumeration; outputs are indexed by cell and mode only) ----------------------------
    struct Edge {
        int64_t key;
        int32_t cell, loc;
    };
    std::vector<Edge> edges((size_t)(3 * ncells));
    for (int64_t c = 0; c < ncells; ++c)
        for (int l = 0; l < 3; ++l) {
            int64_t a = ctx->h_cellnodes[3 * c + l], b = ctx->h_cellnodes[3 * c + (l + 1) % 3];
            edges[3 * c + l] = {std::min(a, b) * ctx->nnodes + std::max(a, b), (int32_t)c, (int32_t)l};
        }
    std::sort(edges.begin(), edges.end(), [](const Edge& x, const Edge& y) {
        return x.key != y.key ? x.key < y.key : x.cell < y.cell;
    });
    std::vector<int32_t> face_nodes, face_cells, cell_faces((size_t)(3 * ncells));
    for (size_t k = 0; k < edges.size();) {
        size_t k2 = k + 1;
        while (k2 < edges.size() && edges[k2].key == edges[k].key) ++k2;
        ASG_CHECK(ctx, k2 - k <= 2, ASGFEM_EINVAL, "estimate: a mesh edge belongs to more than two cells");
        int32_t f = (int32_t)(face_nodes.size() / 2);
        int32_t c = edges[k].cell, l = edges[k].loc;
        face_nodes.push_back(ctx->h_cellnodes[3 * c + l]);
        face_nodes.push_back(ctx->h_cellnodes[3 * c + (l + 1) % 3]);
        face_cells.push_back(c);
        face_cells.push_back(k2 - k == 2 ? edges[k + 1].cell : -1);
        for (size_t t = k; t < k2; ++t) cell_faces[3 * edges[t].cell + edges[t].loc] = f;
        k = k2;
    }
    const int64_t nfaces = (int64_t)face_cells.size() / 2;
    const int64_t ldE = (N_ext + 15) / 16 * 16;

    DevBuf d_cptr, d_cm, d_ck, d_cg, d_fn, d_fc, d_cf, d_fq, d_E, d_JF, d_part, d_sums, d_stage;
    auto upload = [&](DevBuf& b, const void* h, size_t bytes) -> int {
        ASG_CUDA(ctx, cudaMalloc(&b.p, std::max<size_t>(bytes, 8)));
        if (bytes) ASG_CUDA(ctx, cudaMemcpyAsync(b.p, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return 0;
    };
    int rc = 0;
    if ((rc = upload(d_cptr, cptr.data(), cptr.size() * 4))) return rc;
    if ((rc = upload(d_cm, cm.data(), cm.size() * 4))) return rc;
    if ((rc = upload(d_ck, ck.data(), ck.size() * 4))) return rc;
    if ((rc = upload(d_cg, cg.data(), cg.size() * 8))) return rc;
    if ((rc = upload(d_fn, face_nodes.data(), face_nodes.size() * 4))) return rc;
    if ((rc = upload(d_fc, face_cells.data(), face_cells.size() * 4))) return rc;
    if ((rc = upload(d_cf, cell_faces.data(), cell_faces.size() * 4))) return rc;
    if (f_at_qp && (rc = upload(d_fq, f_at_qp, sizeof(double) * nq * ncells))) return rc;
    ASG_CUDA(ctx, cudaMalloc(&d_E.p, sizeof(double) * ncells * ldE));
    ASG_CUDA(ctx, cudaMalloc(&d_JF.p, sizeof(double) * nfaces * ldE));
    ASG_CUDA(ctx, cudaMemcpyToSymbolAsync(e_xref, xref, sizeof(double) * 2 * nq, 0, cudaMemcpyHostToDevice, ctx->stream));
    ASG_CUDA(ctx, cudaMemcpyToSymbolAsync(e_w, w, sizeof(double) * nq, 0, cudaMemcpyHostToDevice, ctx->stream));
    ASG_CUDA(ctx, cudaMemcpyToSymbolAsync(e_sf, sf, sizeof(double) * nqf, 0, cudaMemcpyHostToDevice, ctx->stream));
    ASG_CUDA(ctx, cudaMemcpyToSymbolAsync(e_wf, wf, sizeof(double) * nqf, 0, cudaMemcpyHostToDevice, ctx->stream));

    EstArgs a;
    a.ld = ctx->ld;
    a.pos = ctx->d_pos;
    a.ldE = ldE;
    a.ncells = ncells;
    a.nfaces = nfaces;
    a.N = (int)N;
    a.N_ext = (int)N_ext;
    a.M_ext = (int)M_ext;
    a.order = ctx->order;
    a.nd = ctx->ndofs4cell;
    a.nq = nq;
    a.nqf = nqf;
    a.u = u;
    a.coords = ctx->d_coords;
    a.cellnodes = ctx->d_cellnodes;
    a.celldofs = ctx->d_celldofs;
    a.cptr = (const int32_t*)d_cptr.p;
    a.cm = (const int32_t*)d_cm.p;
    a.ck = (const int32_t*)d_ck.p;
    a.cg = (const double*)d_cg.p;
    a.fq = f_at_qp ? (const double*)d_fq.p : nullptr;
    a.mean = ctx->mean;
    a.decay = ctx->d_decay;
    a.b1 = ctx->d_b1;
    a.b2 = ctx->d_b2;
    a.face_nodes = (const int32_t*)d_fn.p;
    a.face_cells = (const int32_t*)d_fc.p;
    a.face_loc = nullptr;
    a.cell_faces = (const int32_t*)d_cf.p;
    a.E = (double*)d_E.p;
    a.JF = (double*)d_JF.p;

    size_t smem_vol = sizeof(double) * ((size_t)N + (size_t)nq * (M_ext + 1));
    size_t smem_jmp = sizeof(double) * ((size_t)N * 4 + (size_t)nqf * (M_ext + 1));
    ASG_CHECK(ctx, smem_vol <= 200 * 1024 && smem_jmp <= 200 * 1024, ASGFEM_EINVAL,
              "estimate: too many active modes for the shared-memory tables (N <= 6000 supported)");
    const int gridc = (int)std::min<int64_t>(ncells, 148 * 8), gridf = (int)std::min<int64_t>(nfaces, 148 * 8);
    DevBuf d_mi, d_den, d_lam, d_z1, d_Z2;
    if (kind == 0) {
        ASG_CUDA(ctx, cudaFuncSetAttribute(k_est_volume, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ASG_CUDA(ctx, cudaFuncSetAttribute(k_est_jumps, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ASG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        k_est_volume<<<gridc, 256, smem_vol, ctx->stream>>>(a);
        k_est_jumps<<<gridf, 256, smem_jmp, ctx->stream>>>(a);
        ASG_CUDA(ctx, cudaGetLastError());
    } else {
        // log-transformed primal problem
        ASG_CHECK(ctx, N_ext <= 2048, ASGFEM_EINVAL, "estimate (log-primal): at most 2048 extended multi-indices");
        std::vector<int32_t> mi32((size_t)(N_ext * M_ext));
        std::vector<double> den((size_t)N_ext);
        for (int64_t j = 0; j < N_ext; ++j) {
            double fac = 1.0;
            int64_t deg = 0;
            for (int64_t d = 0; d < M_ext; ++d) {
                const int64_t e = mi_ext[j * M_ext + d];
                mi32[(size_t)(j * M_ext + d)] = (int32_t)e;
                for (int64_t r = 2; r <= e; ++r) fac *= (double)r;
                deg += e;
            }
            den[(size_t)j] = std::sqrt(fac) * ((deg & 1) ? -1.0 : 1.0);
        }
        if ((rc = upload(d_mi, mi32.data(), mi32.size() * 4))) return rc;
        if ((rc = upload(d_den, den.data(), den.size() * 8))) return rc;
        if (lam_at_qp && (rc = upload(d_lam, lam_at_qp, sizeof(double) * (size_t)N_ext * nq * ncells))) return rc;
        ASG_CUDA(ctx, cudaMalloc(&d_z1.p, sizeof(double) * (size_t)ncells));
        ASG_CUDA(ctx, cudaMalloc(&d_Z2.p, sizeof(double) * (size_t)ncells * ldE));
        EstLogArgs L;
        L.mi = (const int32_t*)d_mi.p;
        L.den = (const double*)d_den.p;
        L.lam_qp = lam_at_qp ? (const double*)d_lam.p : nullptr;
        L.ntrunc = ntrunc;
        L.zeta1 = (double*)d_z1.p;
        L.Z2 = (double*)d_Z2.p;
        const size_t smem_log = sizeof(double) * (3 * (size_t)N + (size_t)std::max<int64_t>(ntrunc, M_ext) + 1 + 2 * ((size_t)M_ext + 1));
        ASG_CHECK(ctx, smem_log <= 200 * 1024, ASGFEM_EINVAL, "estimate (log-primal): too many active modes for the shared-memory tables");
        ASG_CUDA(ctx, cudaFuncSetAttribute(k_estlog_volume, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ASG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        k_estlog_volume<<<gridc, 256, smem_log, ctx->stream>>>(a, L);
        k_estlog_jumps<<<gridf, 256, 0, ctx->stream>>>(a);
        ASG_CUDA(ctx, cudaGetLastError());
    }

    // column sums: volume part over cells, jump part over faces (each interior face once, :414)
    const int chunk = 256;
    const int ncc = (int)((ncells + chunk - 1) / chunk), nfc = (int)((nfaces + chunk - 1) / chunk);
    ASG_CUDA(ctx, cudaMalloc(&d_part.p, sizeof(double) * (size_t)std::max(ncc, nfc) * N_ext));
    ASG_CUDA(ctx, cudaMalloc(&d_sums.p, sizeof(double) * (3 * N_ext + 1)));  // volume, jumps, [log-primal: zeta2 per mode, zeta1]
    ASG_CUDA(ctx, cudaMemsetAsync(d_sums.p, 0, sizeof(double) * (3 * N_ext + 1), ctx->stream));
    dim3 gb(ncc, (unsigned)((N_ext + 127) / 128));
    // row-sharded run (asgfem_set_owned_cells): a cell counts where it is owned, an interior face with the share of its two
    // cells that this rank owns (1, 1/2 or 0: the neighbour adds the other half); the sums are all-reduced below
    DevBuf d_wc, d_wf;
    const double *wcell = nullptr, *wface = nullptr;
    if (!ctx->h_cell_owned.empty()) {
        ASG_CHECK(ctx, (int64_t)ctx->h_cell_owned.size() == ncells, ASGFEM_ESTATE, "estimate: owned-cell flags do not match the mesh");
        std::vector<double> wc((size_t)ncells), wfv((size_t)nfaces);
        for (int64_t c = 0; c < ncells; ++c) wc[(size_t)c] = ctx->h_cell_owned[(size_t)c] ? 1.0 : 0.0;
        for (int64_t f = 0; f < nfaces; ++f) {
            const int32_t c0 = face_cells[(size_t)(2 * f)], c1 = face_cells[(size_t)(2 * f + 1)];
            wfv[(size_t)f] = 0.5 * ((c0 >= 0 ? wc[(size_t)c0] : 0.0) + (c1 >= 0 ? wc[(size_t)c1] : 0.0));
        }
        if ((rc = upload(d_wc, wc.data(), sizeof(double) * ncells))) return rc;
        if ((rc = upload(d_wf, wfv.data(), sizeof(double) * nfaces))) return rc;
        wcell = (const double*)d_wc.p, wface = (const double*)d_wf.p;
    }
    k_colsum_partial<<<gb, 128, 0, ctx->stream>>>(a.E, ncells, ldE, (int)N_ext, chunk, (double*)d_part.p, wcell);
    k_colsum_final<<<(unsigned)((N_ext + 127) / 128), 128, 0, ctx->stream>>>((double*)d_part.p, ncc, (int)N_ext, (double*)d_sums.p);
    dim3 gf(nfc, (unsigned)((N_ext + 127) / 128));
    k_colsum_partial<<<gf, 128, 0, ctx->stream>>>(a.JF, nfaces, ldE, (int)N_ext, chunk, (double*)d_part.p, wface);
    k_colsum_final<<<(unsigned)((N_ext + 127) / 128), 128, 0, ctx->stream>>>((double*)d_part.p, nfc, (int)N_ext,
                                                                            (double*)d_sums.p + N_ext);
    if (kind == 1) {  // data terms: zeta2 per mode over the cells, zeta1 over the cells
        k_colsum_partial<<<gb, 128, 0, ctx->stream>>>((const double*)d_Z2.p, ncells, ldE, (int)N_ext, chunk, (double*)d_part.p, wcell);
        k_colsum_final<<<(unsigned)((N_ext + 127) / 128), 128, 0, ctx->stream>>>((double*)d_part.p, ncc, (int)N_ext,
                                                                                (double*)d_sums.p + 2 * N_ext);
        k_colsum_partial<<<dim3(ncc, 1), 128, 0, ctx->stream>>>((const double*)d_z1.p, ncells, 1, 1, chunk, (double*)d_part.p, wcell);
        k_colsum_final<<<1, 128, 0, ctx->stream>>>((double*)d_part.p, ncc, 1, (double*)d_sums.p + 3 * N_ext);
    }
    k_add_face_jumps<<<gridc, 256, 0, ctx->stream>>>(a.E, a.JF, a.cell_faces, ncells, ldE, (int)N_ext);
    ASG_CUDA(ctx, cudaGetLastError());
    ASG_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    if ((rc = dist_allreduce_sum(ctx, (double*)d_sums.p, (size_t)(3 * N_ext + 1)))) return rc;  // no-op without a communicator
    std::vector<double> sums((size_t)(3 * N_ext + 1));
    ASG_CUDA(ctx, cudaMemcpyAsync(sums.data(), d_sums.p, sizeof(double) * (3 * N_ext + 1), cudaMemcpyDeviceToHost, ctx->stream));

    // marking indicator: row sums over the selected columns (what scripts/poisson.jl:402 takes from eta4cell) - ncells doubles
    // instead of the ncells x N_ext matrix
    DevBuf d_sel, d_cs;
    if (cellsum && nsel >= 0) {
        std::vector<int32_t> sel32((size_t)std::max<int64_t>(nsel, 1), 0);
        for (int64_t k = 0; k < nsel; ++k) {
            ASG_CHECK(ctx, sel[k] >= 1 && sel[k] <= N_ext, ASGFEM_EINVAL, "estimate: selected column out of range");
            sel32[(size_t)k] = (int32_t)(sel[k] - 1);
        }
        if ((rc = upload(d_sel, sel32.data(), sizeof(int32_t) * sel32.size()))) return rc;
        ASG_CUDA(ctx, cudaMalloc(&d_cs.p, sizeof(double) * (size_t)ncells));
        k_rowsum_selected<<<(unsigned)((ncells + 127) / 128), 128, 0, ctx->stream>>>(a.E, ncells, ldE, (const int32_t*)d_sel.p, (int)nsel,
                                                                                  (double*)d_cs.p);
        ASG_CUDA(ctx, cudaMemcpyAsync(cellsum, d_cs.p, sizeof(double) * (size_t)ncells, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (eta4cell) {
        // eta4cell to the host in the Julia layout (ncells x N_ext, column-major), chunk of columns at a time.  The caller's
        // array is pageable; pinning it for the duration of the call turns the copies into direct DMA (large outputs only)
        const size_t out_bytes = sizeof(double) * (size_t)ncells * (size_t)N_ext;
        const bool pinned_out = out_bytes >= (64u << 20) && cudaHostRegister(eta4cell, out_bytes, cudaHostRegisterDefault) == cudaSuccess;
        if (!pinned_out) (void)cudaGetLastError();
        struct Unpin {
            void* p;
            ~Unpin() {
                if (p) cudaHostUnregister(p);
            }
        } unpin{pinned_out ? (void*)eta4cell : nullptr};
        int64_t jc_max = std::max<int64_t>(1, std::min<int64_t>(N_ext, (256ll << 20) / (8 * std::max<int64_t>(ncells, 1))));
        ASG_CUDA(ctx, cudaMalloc(&d_stage.p, sizeof(double) * ncells * jc_max));
        for (int64_t j0 = 0; j0 < N_ext; j0 += jc_max) {
            int jc = (int)std::min<int64_t>(jc_max, N_ext - j0);
            dim3 g((unsigned)((ncells + 31) / 32), (unsigned)((jc + 31) / 32));
            k_cols_to_stage<<<g, dim3(32, 8), 0, ctx->stream>>>(a.E, (double*)d_stage.p, ncells, ldE, j0, jc);
            ASG_CUDA(ctx, cudaMemcpyAsync(eta4cell + ncells * j0, d_stage.p, sizeof(double) * ncells * jc, cudaMemcpyDeviceToHost,
                                          ctx->stream));
        }
    }
    ASG_CUDA(ctx, cudaGetLastError());
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    {
        float ms = 0;
        ASG_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->last_estimate_ms = ms;
    }
    for (int64_t j = 0; j < N_ext; ++j) {
        double vol = std::sqrt(sums[j]);                       // eta4modes[j] = sqrt(sum(eta4cell[:,j]))  (:362-364 / :219-221)
        if (kind == 0)
            eta4modes[j] = std::sqrt(vol * vol + sums[N_ext + j]);  // sqrt(eta4modes[j]^2 + sum(jumps4face))   (:414)
        else  // log-primal (:244): eta4modes[j] += sqrt(eta4modes[j]^2 + sum(jumps4face)) for the active modes - "+=" as in the reference
            eta4modes[j] = j < N ? vol + std::sqrt(vol * vol + sums[N_ext + j]) : vol;
    }
    if (kind == 1 && zeta3) {
        double z2 = 0.0;
        for (int64_t j = 0; j < N_ext; ++j) z2 += sums[2 * N_ext + j];
        zeta3[1] = sums[3 * N_ext];
        zeta3[2] = z2;
        zeta3[0] = zeta3[1] - zeta3[2];  // zeta_data = zeta_data1 - zeta_data2   (:248, :256)
    }
    return 0;
}

}  // namespace asgfem
