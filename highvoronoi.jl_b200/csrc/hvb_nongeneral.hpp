// hvb_nongeneral.hpp -- host side of "non-general position resolved by perturbation + merge" (DESIGN section 13):
// the constants, the deterministic offset of a generator, the merge of rows with equal coordinates into variable-length
// signatures and the neighbour lists of such a mesh.  Plain C++ shared by the library (Ctx::resolve_degenerate, hvb_ctx.cuh)
// and the CPU harness (tests/hostsim), so that the whole pipeline is tested without a GPU as well.
#pragma once
#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <unordered_map>
#include <utility>
#include <vector>
#include "hvb_core.cuh"

#ifndef HVB_PERTURB_REL
#define HVB_PERTURB_REL 1e-9      // offset of a generator / extent of the cloud
#endif
#ifndef HVB_MERGE_REL
#define HVB_MERGE_REL 1e-8        // rows closer than this (times the extent, per coordinate) are one vertex
#endif
#ifndef HVB_FLAT_TOL
#define HVB_FLAT_TOL 1e-7         // |det| of the unit edge vectors of a simplex below which its generators count as coplanar
#endif
#ifndef HVB_TMIN_REL
#define HVB_TMIN_REL -1e-13       // smallest accepted ray parameter / extent on a perturbed cloud (Dev::t_min)
#endif

namespace hvb {

// offset of coordinate `flat_index` (= generator * dim + axis, caller order) in units of the amplitude: uniform in (-1, 1),
// a function of the index alone (k_perturb on the device, the host harness)
HVB_HD double perturb_unit(u64 flat_index) {
    const u64 h = mix64(flat_index * 0x9e3779b97f4a7c15ULL + 0x243f6a8885a308d3ULL);
    return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

// Rows (sorted 1-based ids [nrow][dim+1], coordinates [nrow][dim]) whose coordinates agree to `eps` per axis are one vertex: its
// signature is the union of theirs.  Two grids of cell size eps, shifted by half a cell: the members of a cluster agree to ~1e-13
// of the extent, so they share a cell of at least one grid unless they straddle a boundary of both (probability ~ (d 1e-13 / eps)^2);
// union-find over both.  Output: CSR signatures in lexicographic order (if `sort`), the coordinates of the first (smallest) row of
// every cluster, the largest signature length and the number of vertices with more than dim + 1 generators.
inline void merge_rows(int dim, int64_t nrow, const int64_t* sig, const double* r, double eps, bool sort,
                       std::vector<int64_t>& off, std::vector<int64_t>& ids, std::vector<double>& rout, int64_t& maxlen, int64_t& ndegenerate) {
    std::vector<int> parent((size_t)nrow);
    for (int64_t i = 0; i < nrow; ++i) parent[i] = (int)i;
    auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    auto unite = [&](int a, int b) { a = find(a); b = find(b); if (a != b) { if (a < b) parent[b] = a; else parent[a] = b; } };
    for (int pass = 0; pass < 2; ++pass) {
        std::unordered_map<u64, int> first;
        first.reserve((size_t)nrow * 2);
        for (int64_t i = 0; i < nrow; ++i) {
            u64 h = 0x9ae16a3b2f90404fULL + (u64)pass;
            for (int k = 0; k < dim; ++k) {
                const long long c = (long long)floor(r[i * dim + k] / eps + 0.5 * pass);
                h = mix64(h ^ ((u64)c + 0x9e3779b97f4a7c15ULL * (u64)(k + 1)));
            }
            auto it = first.find(h);
            if (it == first.end()) { first.emplace(h, (int)i); continue; }
            const int j = it->second;
            double dmax = 0;
            for (int k = 0; k < dim; ++k) dmax = std::max(dmax, fabs(r[i * dim + k] - r[(size_t)j * dim + k]));
            if (dmax <= eps) unite((int)i, j);
        }
    }
    // clusters in the order of their first row (rows are sorted by signature: the first row is the smallest one)
    std::vector<int> cluster_of((size_t)nrow, -1), head;
    for (int64_t i = 0; i < nrow; ++i) {
        const int rt = find((int)i);
        if (cluster_of[rt] < 0) { cluster_of[rt] = (int)head.size(); head.push_back(rt); }
        cluster_of[i] = cluster_of[rt];
    }
    const size_t nc = head.size();
    std::vector<std::vector<int64_t> > sets(nc);
    for (int64_t i = 0; i < nrow; ++i) { auto& v = sets[cluster_of[i]]; v.insert(v.end(), sig + i * (dim + 1), sig + (i + 1) * (dim + 1)); }
    maxlen = dim + 1; ndegenerate = 0;
    for (auto& v : sets) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        maxlen = std::max<int64_t>(maxlen, (int64_t)v.size());
        if ((int64_t)v.size() > dim + 1) ++ndegenerate;
    }
    std::vector<int> order(nc);
    for (size_t c = 0; c < nc; ++c) order[c] = (int)c;
    if (sort) std::sort(order.begin(), order.end(), [&](int a, int b) { return sets[a] < sets[b]; });
    off.assign(nc + 1, 0); ids.clear(); rout.resize(nc * dim);
    for (size_t o = 0; o < nc; ++o) {
        const int c = order[o];
        ids.insert(ids.end(), sets[c].begin(), sets[c].end());
        off[o + 1] = (int64_t)ids.size();
        for (int k = 0; k < dim; ++k) rout[o * dim + k] = r[(size_t)head[c] * dim + k];
    }
}

// Neighbour lists of a merged mesh: i and j are neighbours if they share a FULL interface (neighbors.jl:205-212; the reference's
// NeighborFinder removes cells that only share a lower-dimensional face of a non-general vertex): the vertices (and unbounded
// edges: ray_edge [nrays][dim] ids, ray_dir [nrays][dim]) both belong to span an affine space of dimension dim - 1.
// CSR over the n generators (1-based ids; planes n + p appear as neighbours, have no list), ids ascending.
inline void merged_neighbors(int dim, int64_t n, int64_t nvert, const int64_t* off, const int64_t* ids, const double* r,
                             int64_t nrays, const int64_t* ray_edge, const double* ray_dir,
                             std::vector<int64_t>& nb_off, std::vector<int64_t>& nb_ids) {
    struct Item { int64_t i, j; int64_t v; };            // v >= 0: merged vertex; v < 0: unbounded edge -v - 1
    std::vector<Item> items;
    for (int64_t v = 0; v < nvert; ++v)
        for (int64_t a = off[v]; a < off[v + 1]; ++a)
            for (int64_t b = a + 1; b < off[v + 1]; ++b) items.push_back({ids[a], ids[b], v});
    for (int64_t q = 0; q < nrays; ++q)
        for (int a = 0; a < dim; ++a)
            for (int b = a + 1; b < dim; ++b) items.push_back({ray_edge[q * dim + a], ray_edge[q * dim + b], -q - 1});
    std::sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.i != y.i ? x.i < y.i : (x.j != y.j ? x.j < y.j : x.v < y.v); });
    std::vector<std::pair<int64_t, int64_t> > adj;      // (cell, neighbour)
    const double tol = 1e-6;
    for (size_t a = 0; a < items.size();) {
        size_t b = a;
        while (b < items.size() && items[b].i == items[a].i && items[b].j == items[a].j) ++b;
        // rank of the span: Gram-Schmidt over the differences to the first vertex and the directions of the unbounded edges
        double basis[6][6];
        int rank = 0;
        const double* p0 = nullptr;
        double scale = 0;
        for (size_t t = a; t < b && !p0; ++t) if (items[t].v >= 0) p0 = &r[(size_t)items[t].v * dim];
        for (size_t t = a; t < b; ++t)
            if (items[t].v >= 0 && p0) {
                double s2 = 0;
                for (int k = 0; k < dim; ++k) { const double dd = r[(size_t)items[t].v * dim + k] - p0[k]; s2 += dd * dd; }
                scale = std::max(scale, sqrt(s2));
            }
        for (size_t t = a; t < b && rank < dim - 1; ++t) {
            double w[6];
            double ref;
            if (items[t].v >= 0) { if (!p0) continue; for (int k = 0; k < dim; ++k) w[k] = r[(size_t)items[t].v * dim + k] - p0[k]; ref = scale; }
            else { for (int k = 0; k < dim; ++k) w[k] = ray_dir[(size_t)(-items[t].v - 1) * dim + k]; ref = 1.0; }
            for (int rep = 0; rep < 2; ++rep)
                for (int q = 0; q < rank; ++q) {
                    double sdot = 0;
                    for (int k = 0; k < dim; ++k) sdot += w[k] * basis[q][k];
                    for (int k = 0; k < dim; ++k) w[k] -= sdot * basis[q][k];
                }
            double nw = 0;
            for (int k = 0; k < dim; ++k) nw += w[k] * w[k];
            nw = sqrt(nw);
            if (ref > 0 && nw > tol * ref) { for (int k = 0; k < dim; ++k) basis[rank][k] = w[k] / nw; ++rank; }
        }
        if (rank >= dim - 1) {
            const int64_t i = items[a].i, j = items[a].j;
            if (i <= n) adj.emplace_back(i, j);
            if (j <= n) adj.emplace_back(j, i);
        }
        a = b;
    }
    std::sort(adj.begin(), adj.end());
    nb_off.assign((size_t)n + 1, 0); nb_ids.clear(); nb_ids.reserve(adj.size());
    for (auto& pr : adj) { nb_off[pr.first]++; nb_ids.push_back(pr.second); }
    for (int64_t i = 0; i < n; ++i) nb_off[i + 1] += nb_off[i];
}

}  // namespace hvb
