// Host index machinery: multi-index hashing, neighbour tables, coupling weights, G, extended sets.
// Integer/index structures must match the reference bit-exactly (BASELINE.json north_star).
//   get_neighbours                 src/estimate.jl:1-22
//   get_tensor_multiplication_with_ym  src/tensorizedbasis.jl:193-218
//   normalise_recurrence_coefficients  src/orthogonal_polynomials/orthogonal_polynomials.jl:211-221
//   add_boundary_modes             src/mopcontrol.jl:60-132
//   classify_modes                 src/mopcontrol.jl:168-241
// The reference finds neighbours by O(N^2 M) linear scans; here a hash map gives O(N M).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_set>

#include "common.h"

namespace asgfem {

uint64_t hash_mi(const int64_t* v, int64_t M) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int64_t k = 0; k < M; ++k) {
        uint64_t x = (uint64_t)v[k] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        x ^= x >> 30;
        x *= 0xBF58476D1CE4E5B9ull;
        x ^= x >> 27;
        h ^= x;
    }
    return h;
}

namespace {
struct VecHash {
    size_t operator()(const std::vector<int64_t>& v) const { return (size_t)hash_mi(v.data(), (int64_t)v.size()); }
};
using MiMap = std::unordered_map<std::vector<int64_t>, int64_t, VecHash>;
using MiSet = std::unordered_set<std::vector<int64_t>, VecHash>;
}  // namespace

int64_t MultiIndexSet::maxdeg() const {
    int64_t d = 0;
    for (int64_t v : mi) d = std::max(d, v);
    return d;
}

void MultiIndexSet::build_neighbours() {
    MiMap pos;
    pos.reserve((size_t)N * 2);
    std::vector<int64_t> key((size_t)M);
    for (int64_t j = 0; j < N; ++j) {
        key.assign(mi.begin() + j * M, mi.begin() + (j + 1) * M);
        pos[key] = j + 1;  // the reference's full scan keeps the LAST match
    }
    plus.assign((size_t)(M * N), 0);
    minus.assign((size_t)(M * N), 0);
    for (int64_t j = 0; j < N; ++j) {
        key.assign(mi.begin() + j * M, mi.begin() + (j + 1) * M);
        for (int64_t m = 0; m < M; ++m) {
            key[m] += 1;
            auto it = pos.find(key);
            if (it != pos.end()) plus[m + M * j] = it->second;
            key[m] -= 2;
            if (key[m] >= 0) {
                it = pos.find(key);
                if (it != pos.end()) minus[m + M * j] = it->second;
            }
            key[m] += 1;
        }
    }
}

void coupling_weights(int family, int64_t maxdeg, std::vector<double>& gp, std::vector<double>& gm) {
    gp.assign((size_t)maxdeg + 1, 0.0);
    gm.assign((size_t)maxdeg + 1, 0.0);
    for (int64_t k = 0; k <= maxdeg; ++k) {
        if (family == ASGFEM_LEGENDRE) {
            // (a,b,c) = (0, (2k+1)//(k+1), k//(k+1)); norms h_k = sqrt(2/(2k+1)) (Legendre_uniform.jl:20-21,32);
            // normalised b' = b*h1/h2, c' = c*h0/h2, evaluated left to right in Float64 like Julia
            double b = (double)(2 * k + 1) / (double)(k + 1);
            double c = (double)k / (double)(k + 1);
            double h2 = std::sqrt(2.0 / (double)(2 * (k + 1) + 1));
            double h1 = std::sqrt(2.0 / (double)(2 * k + 1));
            double h0 = k > 0 ? std::sqrt(2.0 / (double)(2 * (k - 1) + 1)) : 0.0;
            double bn = b * h1 / h2;
            double cn = c * h0 / h2;
            gp[k] = 1.0 / bn;
            gm[k] = cn / bn;
        } else {
            // Hermite: (0,1,k), norms sqrt(k!) in BigFloat (Hermite_normal.jl:20-21,32): 1/b' = sqrt(k+1) and
            // c'/b' = sqrt(k) up to 2^-250, rounded to Float64 on assignment into G -> correctly rounded sqrt
            gp[k] = std::sqrt((double)(k + 1));
            gm[k] = std::sqrt((double)k);
        }
    }
}

void build_coupling(const MultiIndexSet& S, int family, Coupling& C) {
    std::vector<double> gp, gm;
    coupling_weights(family, S.maxdeg() + 1, gp, gm);
    C.ptr.assign((size_t)S.N + 1, 0);
    C.m.clear();
    C.nu.clear();
    C.g.clear();
    struct E {
        int32_t nu, m;
        double g;
    };
    std::vector<E> tmp;
    for (int64_t j = 0; j < S.N; ++j) {
        tmp.clear();
        for (int64_t m = 0; m < S.M; ++m) {
            int64_t deg = S.mi[j * S.M + m];
            int64_t p = S.plus[m + S.M * j], q = S.minus[m + S.M * j];
            if (p > 0) tmp.push_back({(int32_t)(p - 1), (int32_t)(m + 1), gp[deg]});
            if (q > 0) tmp.push_back({(int32_t)(q - 1), (int32_t)(m + 1), gm[deg]});
        }
        // mul! visits `for nu in 1:N, e in 1:M` (solvers_poisson_primal.jl:110)
        std::sort(tmp.begin(), tmp.end(), [](const E& a, const E& b) { return a.nu != b.nu ? a.nu < b.nu : a.m < b.m; });
        for (auto& e : tmp) {
            C.m.push_back(e.m);
            C.nu.push_back(e.nu);
            C.g.push_back(e.g);
        }
        C.ptr[j + 1] = (int32_t)C.m.size();
    }
}

}  // namespace asgfem

using namespace asgfem;

extern "C" int asgfem_coupling_weights(int32_t family, int64_t maxdeg, double* gplus, double* gminus) try {
    if (maxdeg < 0 || !gplus || !gminus || (family != ASGFEM_LEGENDRE && family != ASGFEM_HERMITE)) return ASGFEM_EINVAL;
    std::vector<double> gp, gm;
    coupling_weights(family, maxdeg, gp, gm);
    std::memcpy(gplus, gp.data(), sizeof(double) * gp.size());
    std::memcpy(gminus, gm.data(), sizeof(double) * gm.size());
    return 0;
}
ASG_BOUNDARY_CATCH(nullptr)

extern "C" int asgfem_add_boundary_modes(int64_t N, int64_t M, const int64_t* mi, int64_t p_extension, int64_t tail1,
                                         int64_t tail2, int64_t* N_ext, int64_t* M_ext, int64_t* out,
                                         int64_t out_capacity) try {
    if (N < 1 || M < 1 || !mi || !N_ext || !M_ext) return ASGFEM_EINVAL;
    // first loop (:65-73): j = 2..N, range of k frozen at loop entry, lowest qualifying nonzero wins
    int64_t last_nonzero = 0, maxdegree1 = 0;
    for (int64_t j = 1; j < N; ++j) {
        const int64_t* mj = mi + j * M;
        maxdegree1 = std::max(maxdegree1, mj[0]);
        int64_t lo = last_nonzero + 1;
        for (int64_t k = M; k >= lo; --k)
            if (mj[k - 1] != 0) last_nonzero = k;
    }
    // prepare_multi_indices!(...; minimal_length = last_nonzero + tail_extension[1]) (:75)
    int64_t L = std::max(M, last_nonzero + tail1);
    std::vector<std::vector<int64_t>> base((size_t)N, std::vector<int64_t>((size_t)L, 0));
    for (int64_t j = 0; j < N; ++j) std::copy(mi + j * M, mi + (j + 1) * M, base[j].begin());
    std::vector<std::vector<int64_t>> ext = base;
    MiSet have;
    for (auto& v : ext) have.insert(v);
    auto push = [&](const std::vector<int64_t>& v) {
        if (have.insert(v).second) ext.push_back(v);
    };
    std::vector<int64_t> nw;
    for (int64_t k = 1; k <= L; ++k) {  // :80-87
        nw = base[0];
        nw[k - 1] = 1;
        push(nw);
    }
    for (int64_t k = maxdegree1 + 1; k <= maxdegree1 + p_extension; ++k) {  // :90-97
        nw = base[0];
        nw[0] = k;
        push(nw);
    }
    for (int64_t j = 0; j < N; ++j) {  // :100-130
        int64_t last_nonzero_pos = 1;
        for (int64_t k = L; k >= 1; --k)
            if (base[j][k - 1] != 0) {
                last_nonzero_pos = k;
                break;
            }
        for (int64_t k = 1; k <= last_nonzero_pos + tail2; ++k) {
            if (k > last_nonzero + tail2) break;
            if (k > L) return ASGFEM_EINVAL;  // the Julia original would throw a BoundsError here
            nw = base[j];
            nw[k - 1] += 1;
            push(nw);
        }
    }
    *N_ext = (int64_t)ext.size();
    *M_ext = L;
    if (out) {
        if (out_capacity < *N_ext * L) return ASGFEM_EINVAL;
        for (size_t j = 0; j < ext.size(); ++j) std::copy(ext[j].begin(), ext[j].end(), out + j * L);
    }
    return 0;
}
ASG_BOUNDARY_CATCH(nullptr)

extern "C" int asgfem_classify_modes(int64_t N_ext, int64_t M, const int64_t* mi_ext, int64_t N_active, int32_t* cls) try {
    if (N_ext < 1 || M < 1 || !mi_ext || !cls || N_active < 0 || N_active > N_ext) return ASGFEM_EINVAL;
    MiSet act;
    std::vector<int64_t> v((size_t)M);
    for (int64_t j = 0; j < N_active; ++j) {
        v.assign(mi_ext + j * M, mi_ext + (j + 1) * M);
        act.insert(v);
    }
    for (int64_t j = 0; j < N_ext; ++j) {
        const int64_t* mj = mi_ext + j * M;
        int64_t lnp = 1;
        for (int64_t k = M; k >= 1; --k)
            if (mj[k - 1] != 0) {
                lnp = k;
                break;
            }
        v.assign(mj, mj + M);
        if (act.count(v)) {
            if (lnp == M) {
                cls[j] = 3;  // active_bnd (:188-189)
            } else {
                bool active = true;
                for (int64_t k = 1; k <= lnp + 1; ++k) {
                    v.assign(mj, mj + M);
                    v[k - 1] += 1;
                    if (!act.count(v)) {
                        active = false;
                        break;
                    }
                }
                cls[j] = active ? 4 : 3;
            }
        } else {
            int level = 0;
            int64_t kmax = std::min(lnp + 1, M);
            for (int64_t k = 1; k <= kmax && !level; ++k) {
                v.assign(mj, mj + M);
                if (v[k - 1] > 0) {
                    v[k - 1] -= 1;
                    if (act.count(v)) level = 1;
                }
            }
            if (!level) {
                for (int64_t k = 1; k <= kmax && !level; ++k)
                    for (int64_t k2 = 1; k2 <= kmax && !level; ++k2) {
                        v.assign(mj, mj + M);
                        if (v[k - 1] > 0 && v[k2 - 1] > 0) {
                            v[k - 1] -= 1;
                            v[k2 - 1] -= 1;
                            if (act.count(v)) level = 2;
                        }
                    }
            }
            cls[j] = level;  // 0 inactive_else, 1 inactive_bnd, 2 inactive_bnd2
        }
    }
    return 0;
}
ASG_BOUNDARY_CATCH(nullptr)
