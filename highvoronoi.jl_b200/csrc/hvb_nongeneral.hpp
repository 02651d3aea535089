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

// open-addressing table u64 -> int (insert-or-find), sized once: the merge is a rare path but lattices of 10^6 generators are
// a legitimate input (std::unordered_map was most of its time)
struct FlatMap {
    std::vector<u64> key; std::vector<int> val; u64 mask;
    explicit FlatMap(size_t n) { size_t c = 16; while (c < 2 * n + 2) c <<= 1; key.assign(c, ~0ULL); val.assign(c, -1); mask = c - 1; }
    // returns the stored value of `k` (inserting `v` if absent)
    int get_or_put(u64 k, int v) {
        if (k == ~0ULL) k = 0;
        u64 s = mix64(k) & mask;
        for (;;) {
            if (key[s] == ~0ULL) { key[s] = k; val[s] = v; return v; }
            if (key[s] == k) return val[s];
            s = (s + 1) & mask;
        }
    }
};

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
    const double inv = 1.0 / eps;
    for (int pass = 0; pass < 2; ++pass) {
        FlatMap first((size_t)nrow);
        for (int64_t i = 0; i < nrow; ++i) {
            u64 h = 0x9ae16a3b2f90404fULL + (u64)pass;
            for (int k = 0; k < dim; ++k) {
                const long long c = (long long)floor(r[i * dim + k] * inv + 0.5 * pass);
                h = mix64(h ^ ((u64)c + 0x9e3779b97f4a7c15ULL * (u64)(k + 1)));
            }
            const int j = first.get_or_put(h, (int)i);
            if (j == (int)i) continue;
            double dmax = 0;
            for (int k = 0; k < dim; ++k) dmax = std::max(dmax, fabs(r[i * dim + k] - r[(size_t)j * dim + k]));
            if (dmax <= eps) unite((int)i, j);
        }
    }
    // clusters in the order of their first row (rows are sorted by signature: the first row is the smallest one); rows grouped
    // by cluster with a counting sort
    std::vector<int> cluster_of((size_t)nrow), head;
    for (int64_t i = 0; i < nrow; ++i) {
        const int rt = find((int)i);
        if (rt == (int)i) { cluster_of[i] = (int)head.size(); head.push_back(rt); }     // the root is the smallest member: met first
        else cluster_of[i] = cluster_of[rt];
    }
    const size_t nc = head.size();
    std::vector<int64_t> cstart(nc + 1, 0);
    for (int64_t i = 0; i < nrow; ++i) cstart[cluster_of[i] + 1]++;
    for (size_t c = 0; c < nc; ++c) cstart[c + 1] += cstart[c];
    std::vector<int> member((size_t)nrow);
    { std::vector<int64_t> cur(cstart.begin(), cstart.end() - 1); for (int64_t i = 0; i < nrow; ++i) member[cur[cluster_of[i]]++] = (int)i; }
    // union of the signatures of every cluster
    std::vector<int64_t> uoff(nc + 1, 0), uids;
    uids.reserve((size_t)nrow * (dim + 1) / 2 + 16);
    std::vector<int64_t> buf;
    maxlen = dim + 1; ndegenerate = 0;
    for (size_t c = 0; c < nc; ++c) {
        buf.clear();
        for (int64_t t = cstart[c]; t < cstart[c + 1]; ++t) { const int64_t* row = sig + (size_t)member[t] * (dim + 1); buf.insert(buf.end(), row, row + dim + 1); }
        if (cstart[c + 1] - cstart[c] > 1) { std::sort(buf.begin(), buf.end()); buf.erase(std::unique(buf.begin(), buf.end()), buf.end()); }
        uids.insert(uids.end(), buf.begin(), buf.end());
        uoff[c + 1] = (int64_t)uids.size();
        maxlen = std::max<int64_t>(maxlen, (int64_t)buf.size());
        if ((int64_t)buf.size() > dim + 1) ++ndegenerate;
    }
    std::vector<int> order(nc);
    for (size_t c = 0; c < nc; ++c) order[c] = (int)c;
    if (sort) std::sort(order.begin(), order.end(), [&](int a, int b) {
        return std::lexicographical_compare(uids.begin() + uoff[a], uids.begin() + uoff[a + 1], uids.begin() + uoff[b], uids.begin() + uoff[b + 1]); });
    off.assign(nc + 1, 0); ids.clear(); ids.reserve(uids.size()); rout.resize(nc * dim);
    for (size_t o = 0; o < nc; ++o) {
        const int c = order[o];
        ids.insert(ids.end(), uids.begin() + uoff[c], uids.begin() + uoff[c + 1]);
        off[o + 1] = (int64_t)ids.size();
        for (int k = 0; k < dim; ++k) rout[o * dim + k] = r[(size_t)head[c] * dim + k];
    }
}

// Neighbour lists of a merged mesh: i and j are neighbours if they share a FULL interface (neighbors.jl:205-212; the reference's
// NeighborFinder removes cells that only share a lower-dimensional face of a non-general vertex): the vertices (and unbounded
// edges: ray_edge [nrays][dim] ids, ray_dir [nrays][dim]) both belong to span an affine space of dimension dim - 1.
// Cell by cell: the vertices (and rays) of a cell are gathered by a counting sort, its candidate neighbours are the other ids of
// those vertices, sorted locally; per candidate the rank of the shared vertices' differences (Gram-Schmidt; a vector is new if what
// is left of it exceeds 1e-6 of the largest difference) decides.
// CSR over the n generators (1-based ids; planes n + p appear as neighbours, have no list), ids ascending.
inline void merged_neighbors(int dim, int64_t n, int64_t nvert, const int64_t* off, const int64_t* ids, const double* r,
                             int64_t nrays, const int64_t* ray_edge, const double* ray_dir,
                             std::vector<int64_t>& nb_off, std::vector<int64_t>& nb_ids) {
    // cell -> its vertices (v >= 0) and unbounded edges (-q - 1)
    std::vector<int64_t> cstart((size_t)n + 2, 0);
    for (int64_t v = 0; v < nvert; ++v) for (int64_t a = off[v]; a < off[v + 1]; ++a) if (ids[a] <= n) cstart[ids[a] + 1]++;
    for (int64_t q = 0; q < nrays; ++q) for (int a = 0; a < dim; ++a) if (ray_edge[q * dim + a] <= n) cstart[ray_edge[q * dim + a] + 1]++;
    for (int64_t i = 0; i <= n; ++i) cstart[i + 1] += cstart[i];
    std::vector<int64_t> items((size_t)cstart[n + 1]);
    {
        std::vector<int64_t> cur(cstart.begin(), cstart.end() - 1);
        for (int64_t v = 0; v < nvert; ++v) for (int64_t a = off[v]; a < off[v + 1]; ++a) if (ids[a] <= n) items[cur[ids[a]]++] = v;
        for (int64_t q = 0; q < nrays; ++q) for (int a = 0; a < dim; ++a) if (ray_edge[q * dim + a] <= n) items[cur[ray_edge[q * dim + a]]++] = -q - 1;
    }
    const double tol = 1e-6;
    const int full = dim - 1;
    nb_off.assign((size_t)n + 1, 0);
    nb_ids.clear();
    std::vector<std::pair<int64_t, int64_t> > cand;           // (other id, vertex or ray) of the current cell
    for (int64_t i = 1; i <= n; ++i) {
        cand.clear();
        for (int64_t t = cstart[i]; t < cstart[i + 1]; ++t) {
            const int64_t v = items[t];
            if (v >= 0) { for (int64_t a = off[v]; a < off[v + 1]; ++a) if (ids[a] != i) cand.emplace_back(ids[a], v); }
            else { const int64_t q = -v - 1; for (int a = 0; a < dim; ++a) if (ray_edge[q * dim + a] != i) cand.emplace_back(ray_edge[q * dim + a], v); }
        }
        std::sort(cand.begin(), cand.end());
        for (size_t a = 0; a < cand.size();) {
            size_t b = a;
            while (b < cand.size() && cand[b].first == cand[a].first) ++b;
            double basis[6][6];
            int rank = 0;
            const double* p0 = nullptr;
            double scale = 0;
            for (size_t t = a; t < b; ++t)
                if (cand[t].second >= 0) {
                    const double* pv = &r[(size_t)cand[t].second * dim];
                    if (!p0) { p0 = pv; continue; }
                    double s2 = 0;
                    for (int k = 0; k < dim; ++k) { const double dd = pv[k] - p0[k]; s2 += dd * dd; }
                    scale = std::max(scale, sqrt(s2));
                }
            for (size_t t = a; t < b && rank < full; ++t) {
                double w[6];
                double ref;
                if (cand[t].second >= 0) { if (!p0) continue; for (int k = 0; k < dim; ++k) w[k] = r[(size_t)cand[t].second * dim + k] - p0[k]; ref = scale; }
                else { for (int k = 0; k < dim; ++k) w[k] = ray_dir[(size_t)(-cand[t].second - 1) * dim + k]; ref = 1.0; }
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
            if (rank >= full) nb_ids.push_back(cand[a].first);
            a = b;
        }
        nb_off[i] = (int64_t)nb_ids.size();
    }
}

}  // namespace hvb
