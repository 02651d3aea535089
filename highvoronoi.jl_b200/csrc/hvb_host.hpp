// hvb_host.hpp -- host-side helpers shared by the library (hvb_api.cu) and the CPU debugging harness
// (tests/hostsim): grid sizing and capacity estimates.  No device code here.
#pragma once
#include <math.h>
#include <stdint.h>
#include <algorithm>
#include "hvb_core.cuh"

namespace hvb {

// expected Delaunay simplices per point in general position, lowerbound(d,d) (edgeiteratebase.jl:151-153)
inline double simplices_per_point(int d) {
    static const double s[7] = {0, 0, 2.0, 6.77, 31.8, 186.7, 1296.4};
    return s[d];
}
// measured on B200 (profiles/r1_tile_occupancy_sweep.md, knob sweeps): longer rows amortise the per-row geometry
inline int default_points_per_cell(int d) {
    static const int p[7] = {0, 0, 4, 5, 4, 4, 4};
    return p[d];
}
// first probe ball radius / circumradius of the origin vertex (re-tuned on the k_walk_coop kernel, profiles/r1_sweeps_session3.md)
inline double default_probe_scale(int d) {
    static const double s[7] = {0, 0, 1.7, 1.5, 1.3, 1.2, 1.2};
    return s[d];
}
inline int tile_size_for_dim(int d) {
    static const int g[7] = {0, 0, 4, 8, 16, 32, 32};
    return g[d];
}
inline uint64_t next_pow2(uint64_t x) {
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

// Uniform grid over the bounding box [blo, bhi] of the generators with about `ppc` points per cell.
template <int D>
inline int64_t setup_grid(Dev<D>& dv, const double* blo, const double* bhi, int64_t n, int ppc) {
    double ext[D], vol = 1.0, emax = 0.0, diag2 = 0.0;
    for (int k = 0; k < D; ++k) {
        ext[k] = bhi[k] - blo[k];
        emax = std::max(emax, ext[k]);
        diag2 += ext[k] * ext[k];
    }
    if (!(emax > 0)) emax = 1.0;
    for (int k = 0; k < D; ++k) { if (ext[k] < 1e-9 * emax) ext[k] = 1e-9 * emax; vol *= ext[k]; }
    double target = std::max(1.0, (double)n / (double)std::max(1, ppc));
    double hh = pow(vol / target, 1.0 / D);
    int64_t cells = 1;
    for (int k = 0; k < D; ++k) {
        int gk = (int)std::min(4096.0, std::max(1.0, floor(ext[k] / hh + 0.5)));
        dv.g[k] = gk;
        cells *= gk;
    }
    while (cells > (int64_t)1 << 28) {            // keep cell_start addressable and small
        cells = 1;
        for (int k = 0; k < D; ++k) { dv.g[k] = std::max(1, dv.g[k] / 2); cells *= dv.g[k]; }
    }
    dv.hmin = 1e300;
    for (int k = 0; k < D; ++k) {
        dv.lo[k] = blo[k];
        dv.h[k] = ext[k] / dv.g[k];
        dv.inv_h[k] = 1.0 / dv.h[k];
        dv.hmin = std::min(dv.hmin, dv.h[k]);
        dv.h32[k] = (float)dv.h[k];
        dv.inv_h32[k] = (float)dv.inv_h[k];
    }
    dv.ext = emax;
    dv.ext32 = (float)emax;
    dv.diag = sqrt(diag2) > 0 ? sqrt(diag2) : 1.0;
    return cells;
}

inline int64_t estimate_vertices(int d, int64_t n, int nplanes) {
    double v = simplices_per_point(d) * (double)n * 1.25 + 4096.0 + 64.0 * nplanes;
    return (int64_t)v;
}

}  // namespace hvb
