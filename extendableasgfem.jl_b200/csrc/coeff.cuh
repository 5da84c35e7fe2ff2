// Device evaluation of the cosinus KLE modes (get_am!, src/coefficients/cosinus.jl:58-65) and the P1/P2
// reference basis (ExtendableFEMBase H1Pk{1,2,order}), shared by assemble.cu and estimate.cu.
#pragma once
#include <stdint.h>

namespace asgfem {

__device__ __forceinline__ double eval_am(int m, double x, double y, double mean, const double* __restrict__ decay,
                                          const int32_t* __restrict__ b1, const int32_t* __restrict__ b2) {
    if (m == 0) return mean;
    // decay_factors[m] * cos(pi * b1[m] * x[1]) * cos(pi * b2[m] * x[2]), evaluated left to right (cosinus.jl:62)
    return decay[m - 1] * cos(3.141592653589793 * (double)b1[m - 1] * x) * cos(3.141592653589793 * (double)b2[m - 1] * y);
}

// grad a_m (get_gradam!, src/coefficients/cosinus.jl:67-76): the two components are formed first, then scaled by the decay factor
__device__ __forceinline__ void eval_gradam(int m, double x, double y, const double* __restrict__ decay, const int32_t* __restrict__ b1,
                                            const int32_t* __restrict__ b2, double& gx, double& gy) {
    if (m == 0) {
        gx = gy = 0.0;
        return;
    }
    const double pi = 3.141592653589793, c1 = (double)b1[m - 1], c2 = (double)b2[m - 1];
    gx = -c1 * pi * sin(c1 * pi * x) * cos(c2 * pi * y) * decay[m - 1];
    gy = -c2 * pi * cos(c1 * pi * x) * sin(c2 * pi * y) * decay[m - 1];
}

// value of basis function d of the P1 / P2 reference basis at barycentrics lam
template <int ORDER>
__device__ __forceinline__ double basis_value(const double* lam, int d) {
    if (ORDER == 1) return lam[d];
    if (d < 3) return lam[d] * (2.0 * lam[d] - 1.0);
    const int i = d - 3, j = (d - 2) % 3;
    return 4.0 * lam[i] * lam[j];
}

// d(phi_d)/d(lambda_l) of the P2 basis (l_i(2l_i-1), 4 l_i l_j on faces (1,2),(2,3),(3,1)) at barycentrics lam
__device__ __forceinline__ void p2_dphi(const double* lam, int d, double* out3) {
    out3[0] = out3[1] = out3[2] = 0.0;
    if (d < 3) {
        out3[d] = 4.0 * lam[d] - 1.0;
    } else {
        int i = d - 3, j = (d - 2) % 3;
        out3[i] = 4.0 * lam[j];
        out3[j] = 4.0 * lam[i];
    }
}

// gradients of the barycentric coordinates of a triangle and its signed determinant
__device__ __forceinline__ double lambda_gradients(const double* __restrict__ coords, const int32_t* cn, double gl[3][2]) {
    double x1 = coords[2 * cn[0]], y1 = coords[2 * cn[0] + 1];
    double x2 = coords[2 * cn[1]], y2 = coords[2 * cn[1] + 1];
    double x3 = coords[2 * cn[2]], y3 = coords[2 * cn[2] + 1];
    double det = (x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1);
    gl[0][0] = (y2 - y3) / det;
    gl[0][1] = (x3 - x2) / det;
    gl[1][0] = (y3 - y1) / det;
    gl[1][1] = (x1 - x3) / det;
    gl[2][0] = (y1 - y2) / det;
    gl[2][1] = (x2 - x1) / det;
    return det;
}

}  // namespace asgfem
