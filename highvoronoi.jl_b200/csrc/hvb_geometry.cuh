// hvb_geometry.cuh -- geometry products straight from the vertex rows: cell volumes.
//
// Replaces, for general position, what the reference computes downstream of the search with its polygon
// integrators (VI_POLYGON, polyintegrator.jl / integrate.jl:33-53): the volume of every Voronoi cell.  The reference's
// own tests pin the raycast path through exactly this product ("sum of the cell volumes = volume of the domain",
// test/rcmethods.jl:8, test/multithread.jl:8, test/basics.jl:46), so this is also the known-answer test of the search.
//
// Formula (signed orthoscheme / flag decomposition of a simple polytope): a cell P_i of generator x_i is
//   P_i = { y : n_g.y <= b_g }  in coordinates y = x - x_i, with n_g = x_g - x_i, b_g = |n_g|^2 / 2 for a neighbour g
//   and n_p = outward normal, b_p = off_p - n_p.x_i for a boundary plane p.
// For a point c in the affine hull of a face F,  vol_k(F) = 1/k * sum over the facets G of F of h_G(c) * vol_{k-1}(G)
// with the signed height h_G(c) of G over c inside aff(F).  Taking for c the foot point of the previous level and
// unrolling down to the vertices: vol(P_i) = 1/d! * sum over vertices v of P_i, sum over the d! orders in which the d
// facets through v can be imposed, of the product of the d signed heights.  In general position the facets through
// a vertex are the d other entries of its signature, so a vertex row contributes to the d+1 cells (generators) it
// names without any face lattice.  Written once for host and device (tests/hostsim checks it against Qhull).
#pragma once
#include "hvb_core.cuh"

namespace hvb {

// sum over the d! orders of the products of signed heights for the vertex with caller-numbered signature s[0..D]
// (1-based, generators <= n, plane p = n + p) seen from the cell of the generator at position `pos`.
// xs: generators in caller order [n][D].  Divide the sum over all vertices of the cell by d! for the volume.
// `first` (a position of s other than pos, or -1): only the orders that impose that facet first are summed and its own
// height is left out of the products -- the vertex's share of (d-1)! times the (d-1)-volume of the interface between
// the cell and s[first] (VoronoiData(...).area of the reference).
template <int D>
HVB_HD double vertex_flag_sum(const double* xs, long long n, const PlaneSet* ps, const long long* s, int pos, int first = -1) {
    double nrm[D][D], b[D];
    const double* xi = xs + (size_t)(s[pos] - 1) * D;
    int cnt = 0, jfirst = -1;
    for (int k = 0; k < D + 1; ++k) {
        if (k == pos) continue;
        if (k == first) jfirst = cnt;
        const long long g = s[k];
        if (g <= n) {
            const double* xg = xs + (size_t)(g - 1) * D;
            double q = 0;
            for (int a = 0; a < D; ++a) { nrm[cnt][a] = xg[a] - xi[a]; q += nrm[cnt][a] * nrm[cnt][a]; }
            b[cnt] = 0.5 * q;
        } else {
            const int p = (int)(g - n - 1);
            double q = 0;
            for (int a = 0; a < D; ++a) { nrm[cnt][a] = ps->normal[p * 6 + a]; q += nrm[cnt][a] * xi[a]; }
            b[cnt] = ps->off[p] - q;
        }
        ++cnt;
    }
    double Q[D][D], c[D + 1][D], prod[D + 1];
    int it[D];
    unsigned used = 0;
    for (int a = 0; a < D; ++a) c[0][a] = 0.0;
    prod[0] = 1.0;
    int depth = 0;
    it[0] = -1;
    double total = 0.0;
    for (;;) {
        int j = it[depth] + 1;
        if (depth == 0 && jfirst >= 0) j = (it[0] < jfirst) ? jfirst : D;      // one choice at the top level
        while (j < D && ((used >> j) & 1u)) ++j;
        if (j >= D) {                                   // this level is exhausted: back to the previous one
            if (depth == 0) break;
            --depth;
            used &= ~(1u << it[depth]);
            continue;
        }
        it[depth] = j;
        // direction of constraint j inside the current flat: its normal minus the part along the imposed ones (MGS, twice)
        double m[D];
        for (int a = 0; a < D; ++a) m[a] = nrm[j][a];
        for (int rep = 0; rep < 2; ++rep)
            for (int q = 0; q < depth; ++q) {
                double sp = 0;
                for (int a = 0; a < D; ++a) sp += Q[q][a] * m[a];
                for (int a = 0; a < D; ++a) m[a] -= sp * Q[q][a];
            }
        double len2 = 0, nc = 0;
        for (int a = 0; a < D; ++a) { len2 += m[a] * m[a]; nc += nrm[j][a] * c[depth][a]; }
        const double len = sqrt(len2);
        // moving along the unit direction m/len changes n_j.y by m.n_j/len = len: signed height of the facet over the foot
        const double h = (len > 0) ? (b[j] - nc) / len : 0.0;
        const double hf = (depth == 0 && jfirst >= 0) ? 1.0 : h;          // the interface's own height is not part of its area
        if (depth + 1 == D) { total += prod[depth] * hf; continue; }     // a vertex is reached: one flag
        for (int a = 0; a < D; ++a) { Q[depth][a] = m[a] / len; c[depth + 1][a] = c[depth][a] + h * Q[depth][a]; }
        prod[depth + 1] = prod[depth] * hf;
        used |= 1u << j;
        ++depth;
        it[depth] = -1;
    }
    return total;
}

// Integrals of 1, y_a and y_a y_b over the cell, y = x - x_i (the bulk integrals of VoronoiData for polynomial
// integrands up to degree two, integrate.jl:33-53 / polyintegrator.jl -- the reference's own tests integrate x -> [1, x1^2, x2^2],
// test/periodicgrids.jl).  Every flag of the decomposition above is an orthoscheme: the simplex with the vertices
// c_0 = x_i, c_1 (foot on the first facet), ..., c_d (the Voronoi vertex) and signed volume prod(h) / d!.  Over a simplex with
// vertices p_0..p_d:  int y_a = vol * S_a / (d + 1),  int y_a y_b = vol * (S_a S_b + T_ab) / ((d + 1)(d + 2)),  S = sum p_k,
// T_ab = sum p_k,a p_k,b.  The running sums S, T travel down the recursion with the foot points.
// out[0] += sum prod(h);  out[1 + a] += sum prod(h) S_a / (d+1);  out[1 + D + idx(a,b)] += sum prod(h) (S_a S_b + T_ab) / ((d+1)(d+2)),
// idx over a <= b row by row.  Divide by d! for the integrals.
//
// `first` (a position of s other than pos, or -1): the interface between the cell and s[first] instead of the cell -- only the
// orders that impose that facet first, its own height left out: the (d-1)-simplices c_1..c_d inside the facet.
// out[0] += sum prod'(h) (divide by (d-1)! for the area), out[1 + a] += sum prod'(h) S'_a / d with S' = c_1 + ... + c_d
// (the first moment of the interface in local coordinates, (d-1)! times); second moments are not accumulated.
template <int D>
HVB_HD void vertex_flag_moments(const double* xs, long long n, const PlaneSet* ps, const long long* s, int pos, double* out, int first = -1) {
    const int NM = 1 + D + D * (D + 1) / 2;
    for (int a = 0; a < NM; ++a) out[a] = 0.0;
    double nrm[D][D], b[D];
    const double* xi = xs + (size_t)(s[pos] - 1) * D;
    int cnt = 0, jfirst = -1;
    for (int k = 0; k < D + 1; ++k) {
        if (k == pos) continue;
        if (k == first) jfirst = cnt;
        const long long g = s[k];
        if (g <= n) {
            const double* xg = xs + (size_t)(g - 1) * D;
            double q = 0;
            for (int a = 0; a < D; ++a) { nrm[cnt][a] = xg[a] - xi[a]; q += nrm[cnt][a] * nrm[cnt][a]; }
            b[cnt] = 0.5 * q;
        } else {
            const int p = (int)(g - n - 1);
            double q = 0;
            for (int a = 0; a < D; ++a) { nrm[cnt][a] = ps->normal[p * 6 + a]; q += nrm[cnt][a] * xi[a]; }
            b[cnt] = ps->off[p] - q;
        }
        ++cnt;
    }
    double Q[D][D], c[D + 1][D], prod[D + 1];
    double S[D + 1][D], T[D + 1][D * (D + 1) / 2];          // sums over the chain c_0..c_depth (c_0 = 0 contributes nothing)
    int it[D];
    unsigned used = 0;
    for (int a = 0; a < D; ++a) { c[0][a] = 0.0; S[0][a] = 0.0; }
    for (int a = 0; a < D * (D + 1) / 2; ++a) T[0][a] = 0.0;
    prod[0] = 1.0;
    int depth = 0;
    it[0] = -1;
    for (;;) {
        int j = it[depth] + 1;
        if (depth == 0 && jfirst >= 0) j = (it[0] < jfirst) ? jfirst : D;      // one choice at the top level
        while (j < D && ((used >> j) & 1u)) ++j;
        if (j >= D) {
            if (depth == 0) break;
            --depth;
            used &= ~(1u << it[depth]);
            continue;
        }
        it[depth] = j;
        double m[D];
        for (int a = 0; a < D; ++a) m[a] = nrm[j][a];
        for (int rep = 0; rep < 2; ++rep)
            for (int q = 0; q < depth; ++q) {
                double sp = 0;
                for (int a = 0; a < D; ++a) sp += Q[q][a] * m[a];
                for (int a = 0; a < D; ++a) m[a] -= sp * Q[q][a];
            }
        double len2 = 0, nc = 0;
        for (int a = 0; a < D; ++a) { len2 += m[a] * m[a]; nc += nrm[j][a] * c[depth][a]; }
        const double len = sqrt(len2);
        const double h = (len > 0) ? (b[j] - nc) / len : 0.0;
        double cn[D];                                    // the next point of the chain: foot on facet j (the vertex at the last level)
        for (int a = 0; a < D; ++a) cn[a] = c[depth][a] + ((len > 0) ? h * m[a] / len : 0.0);
        const double hf = (depth == 0 && jfirst >= 0) ? 1.0 : h;          // the interface's own height is not part of its area
        if (depth + 1 == D) {
            const double w = prod[depth] * hf;
            out[0] += w;
            double Sa[D];
            // c_0 = 0 adds nothing to S: the same sum serves the d + 1 vertices of the cell's orthoscheme and the d of the facet's
            for (int a = 0; a < D; ++a) { Sa[a] = S[depth][a] + cn[a]; out[1 + a] += w * Sa[a] / (jfirst >= 0 ? D : D + 1); }
            if (jfirst < 0) {
                int q = 0;
                for (int a = 0; a < D; ++a)
                    for (int bb = a; bb < D; ++bb, ++q)
                        out[1 + D + q] += w * (Sa[a] * Sa[bb] + T[depth][q] + cn[a] * cn[bb]) / ((D + 1) * (D + 2));
            }
            continue;
        }
        for (int a = 0; a < D; ++a) { Q[depth][a] = m[a] / len; c[depth + 1][a] = cn[a]; S[depth + 1][a] = S[depth][a] + cn[a]; }
        { int q = 0; for (int a = 0; a < D; ++a) for (int bb = a; bb < D; ++bb, ++q) T[depth + 1][q] = T[depth][q] + cn[a] * cn[bb]; }
        prod[depth + 1] = prod[depth] * hf;
        used |= 1u << j;
        ++depth;
        it[depth] = -1;
    }
}

}  // namespace hvb
