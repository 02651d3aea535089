// hvb_hull.cuh -- convex hull by a facet walk on the min-t query.  Device code that also compiles as host
// C++ (tests/hostsim), like hvb_core.cuh.
//
// Replaces ConvexHull(xs) = systematic_chull (chull.jl:241-387: start at an extreme node, descent_chull, then a queue of
// facets whose sub-facets are explored by raycast_des3, chull.jl:485-499, with an EdgeHashTable counting how often a
// sub-facet was met) WITHOUT computing the interior of the tessellation -- round 1 read the hull off the unbounded edges of
// a complete search.  Not a port: the reference walks an artificial "outer" simplex per facet; here the walk stays on the
// Voronoi diagram and only ever visits its unbounded part:
//
//   * a hull facet (d generators, outward normal u) IS an unbounded Voronoi edge (the same d generators, direction u);
//   * two facets that share a ridge (d-1 generators) are the two unbounded edges of ONE unbounded 2-face of the diagram,
//     the polygon dual to the ridge.  Walking around that polygon from one unbounded edge reaches the other after a few
//     ordinary edge walks: at a vertex R + {p, g} reached over the edge R + {p}, the next edge of the polygon is R + {g}.
//
// So the state of a walk is (vertex, dropped position, pivot position): the edge to walk is the vertex minus the dropped
// generator, and the pivot p is the generator of that edge that does not belong to the ridge being circled.  Each walk is
// the min-t query of the tessellation (min_t_query) -- no new geometric primitive.  A facet table (keyed on the d sorted
// generators) and a ridge table (first facet opens a ridge and sends a walk around it, second facet closes it) are the
// vertex set and the edge table of hvb_core.cuh one dimension down.  The vertices met on the way are not stored in a set:
// they are scratch records of the walk.
#pragma once
#include "hvb_core.cuh"

namespace hvb {

template <int D>
struct HullDev {
    int* fsig;            // [fcap][D] generators of a facet, sorted internal ids
    u32* fitem;           // [fcap] v << 3 | kd : vertex record the unbounded edge starts at, dropped position
    double* fu;           // [fcap][D] outward unit normal
    u32* fcount; u32 fcap;
    u64* ftab; u64 fmask; // facet set: fingerprint << 32 | (index + 1)
    u64* rtab; u64 rmask; // ridge table: [63] valid, [62:36] fingerprint, [35] closed, [34:3] facet index, [2:0] dropped position IN THE FACET
};

// walk entry: [34:6] vertex record, [5:3] dropped position, [2:0] pivot position (7: none)
HVB_HD u64 hull_entry(u32 v, int kd, int kp) { return ((u64)v << 6) | ((u64)kd << 3) | (u64)kp; }

template <int D>
HVB_HD u64 hash_facet(const int* ids, int skip) {
    u64 h = 0x7f4a7c159e3779b9ULL;
#pragma unroll
    for (int i = 0; i < D; ++i)
        if (i != skip) h = (h ^ (u64)(u32)ids[i]) * 0x100000001b3ULL + 0x632be59bd9b4e019ULL;
    return mix64(h);
}

// inserts facet E (D sorted ids) unless present; returns its new index or 0xffffffff
template <int D>
HVB_HD u32 facet_insert(const Dev<D>& dv, const HullDev<D>& hd, const int* E, u32 v, int kd, const double* u) {
    const u64 h = hash_facet<D>(E, -1);
    u64 fp = (h >> 32) << 32;
    if (fp == 0) fp = 1ULL << 32;
    u64 slot = h & hd.fmask;
    u32 mine = 0xffffffffu;
    u64 s = ld_cg(hd.ftab + slot);
    for (;;) {
        if (s == 0) {
            if (mine == 0xffffffffu) {
                mine = atom_add(hd.fcount, 1u);
                if (mine >= hd.fcap) { atom_or(&dv.ctr->flags, (u32)FLAG_RFULL); return 0xffffffffu; }
#pragma unroll
                for (int k = 0; k < D; ++k) { hd.fsig[(size_t)mine * D + k] = E[k]; hd.fu[(size_t)mine * D + k] = u[k]; }
                hd.fitem[mine] = (v << 3) | (u32)kd;
                mem_fence();
            }
            s = atom_cas(hd.ftab + slot, 0ULL, fp | (u64)(mine + 1u));
            if (s == 0) return mine;
        }
        if ((s >> 32) == (fp >> 32)) {
            const int* p = hd.fsig + (size_t)((u32)(s & 0xffffffffu) - 1u) * D;
            bool eq = true;
#pragma unroll
            for (int k = 0; k < D; ++k) eq &= (ld_cg(p + k) == E[k]);
            if (eq) {
                if (mine != 0xffffffffu) hd.fsig[(size_t)mine * D] = -1;          // lost a race: dead record
                return 0xffffffffu;
            }
        }
        slot = (slot + 1) & hd.fmask;
        s = ld_cg(hd.ftab + slot);
    }
}

// is facet E (D sorted ids) stored already?
template <int D>
HVB_HD bool facet_known(const HullDev<D>& hd, const int* E) {
    const u64 h = hash_facet<D>(E, -1);
    u64 fp = (h >> 32) << 32;
    if (fp == 0) fp = 1ULL << 32;
    u64 slot = h & hd.fmask;
    for (;;) {
        const u64 s = ld_cg(hd.ftab + slot);
        if (s == 0) return false;
        if ((s >> 32) == (fp >> 32)) {
            const int* p = hd.fsig + (size_t)((u32)(s & 0xffffffffu) - 1u) * D;
            bool eq = true;
#pragma unroll
            for (int k = 0; k < D; ++k) eq &= (ld_cg(p + k) == E[k]);
            if (eq) return true;
        }
        slot = (slot + 1) & hd.fmask;
    }
}

// registers ridge (E minus position j) of facet f: true = first facet at this ridge (a walk goes around it), false = the
// ridge has both its facets now
template <int D>
HVB_HD bool ridge_register(const HullDev<D>& hd, const int* E, u32 f, int j) {
    const u64 h = hash_facet<D>(E, j);
    const u64 mine = edge_slot(h, f, j);
    u64 slot = h & hd.rmask;
    u64 s = ld_cg(hd.rtab + slot);
    for (;;) {
        if (s == 0) {
            s = atom_cas(hd.rtab + slot, 0ULL, mine);
            if (s == 0) return true;
        }
        if (((s ^ mine) & EDGE_FPMASK) == 0) {
            const u32 f2 = (u32)((s >> 3) & 0xffffffffULL);
            const int j2 = (int)(s & 7);
            const int* p = hd.fsig + (size_t)f2 * D;
            bool eq = true;
            int i2 = 0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                if (i == j) continue;
                if (i2 == j2) ++i2;
                eq &= (ld_cg(p + i2) == E[i]);
                ++i2;
            }
            if (eq) {
                if (!(s & EDGE_CLOSED)) atom_or(hd.rtab + slot, EDGE_CLOSED);
                return false;
            }
        }
        slot = (slot + 1) & hd.rmask;
        s = ld_cg(hd.rtab + slot);
    }
}

HVB_HD void hull_push(u64* q_out, u32* q_count, u32 q_cap, u32* flags, u64 e) {
    const u32 pos = atom_add(q_count, 1u);
    if (pos < q_cap) q_out[pos] = e;
    else atom_or(flags, (u32)FLAG_QFULL);
}

// the unbounded edge (vertex v minus position kd, direction u) is a hull facet: store it, and send a walk around every
// ridge it is the first facet of
template <int D>
HVB_HD void hull_facet_found(const Dev<D>& dv, const HullDev<D>& hd, const int (&sig)[D + 1], u32 v, int kd, const double* u,
                             u64* q_out, u32* q_count, u32 q_cap) {
    int E[D], pos_in_sig[D];
    int c = 0;
#pragma unroll
    for (int k = 0; k < D + 1; ++k) {
        if (k == kd) continue;
        E[c] = sig[k]; pos_in_sig[c] = k; ++c;
    }
    const u32 f = facet_insert<D>(dv, hd, E, v, kd, u);
    if (f == 0xffffffffu) return;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        // circling the ridge E - {E[j]} starts here: at v the polygon continues with the edge v - {E[j]}, whose generator
        // outside the ridge is the one dropped to get the facet
        if (ridge_register<D>(hd, E, f, j)) hull_push(q_out, q_count, q_cap, &dv.ctr->flags, hull_entry(v, pos_in_sig[j], kd));
    }
}

// appends a scratch vertex record (no set: the walk never asks whether it has been here)
template <int D>
HVB_HD u32 hull_new_vertex(const Dev<D>& dv, const int* sig, const double* r) {
    const u32 v = atom_add(dv.vcount, 1u);
    if (v >= dv.vcap) { atom_or(&dv.ctr->flags, (u32)FLAG_VFULL); return 0xffffffffu; }
#pragma unroll
    for (int k = 0; k < D + 1; ++k) dv.vsig[(size_t)v * (D + 1) + k] = sig[k];
#pragma unroll
    for (int k = 0; k < D; ++k) dv.vr[(size_t)v * D + k] = r[k];
    return v;
}

// one step of a walk around a ridge
template <int D, class T>
HVB_HD void hull_step(const Dev<D>& dv, const HullDev<D>& hd, const T& tile, u64 entry, u64* q_out, u32* q_count, u32 q_cap, LocalStats& ls) {
    const u32 v0 = (u32)(entry >> 6);
    const int kd0 = (int)((entry >> 3) & 7), kp = (int)(entry & 7);
    int sig[D + 1];
    RayQ<D> q;
    u32 v; int kd;
    // the lanes of the tile build the same ray and share the rows of the query (the half-space of an unbounded edge spans
    // the whole grid: one lane per ray is hopeless here); lane 0 alone touches the tables
    if (!ray_setup<D>(dv, ((u64)v0 << 3) | (u64)kd0, q, sig, v, kd)) { ls.seed_fail += (tile.lane() == 0); return; }
    {
        // a walk that arrives at a facet another walk has certified already ends here: certifying an unbounded edge means
        // scanning a half-space of the grid, the most expensive query there is, and every facet is reached over several ridges
        int E[D];
#pragma unroll
        for (int j = 0; j < D; ++j) E[j] = (j < kd) ? sig[j] : sig[j + 1];
        if (facet_known<D>(hd, E)) return;
    }
    const Best best = min_t_query<D, T>(dv, tile, q, ls);
    if (tile.lane() != 0) return;
    if (best.id < 0) { hull_facet_found<D>(dv, hd, sig, v, kd, q.u, q_out, q_count, q_cap); return; }
    if (near_tie(best, q.R0sq)) { ls.degenerate++; atom_or(&dv.ctr->flags, (u32)FLAG_DEGEN); }
    // the next vertex of the polygon: edge + winner; the walk goes on over (that vertex minus the pivot)
    int sig2[D + 1];
    int pos = 0;
    {
        int e[D];
#pragma unroll
        for (int j = 0; j < D; ++j) { e[j] = (j < kd) ? sig[j] : sig[j + 1]; pos += (e[j] < best.id) ? 1 : 0; }
#pragma unroll
        for (int i = 0; i < D + 1; ++i) {
            const int lo_ = (i < D) ? e[i] : 0, hi_ = (i > 0) ? e[i - 1] : 0;
            sig2[i] = (i < pos) ? lo_ : ((i == pos) ? best.id : hi_);
        }
    }
    double r2[D];
#pragma unroll
    for (int k = 0; k < D; ++k) r2[k] = q.r[k] + best.t * q.u[k];
    const u32 v2 = hull_new_vertex<D>(dv, sig2, r2);
    if (v2 == 0xffffffffu) return;
    const int pgen = sig[kp];
    int drop2 = 0;
#pragma unroll
    for (int i = 0; i < D + 1; ++i) drop2 = (sig2[i] == pgen) ? i : drop2;
    hull_push(q_out, q_count, q_cap, &dv.ctr->flags, hull_entry(v2, drop2, pos));
}

// the descent of seed_item without the commit: first vertex of the cell of generator `start` (sorted signature, r)
template <int D, class T>
HVB_HD bool descent_vertex(const Dev<D>& dv, const T& tile, int start, int (&ssig)[D + 1], double (&rr)[D], LocalStats& ls) {
    for (int attempt = 0; attempt < 8; ++attempt) {
        int sig[D + 1];
        int cnt = 1;
        sig[0] = start;
        RayQ<D> q;
#pragma unroll
        for (int k = 0; k < D; ++k) { q.x0[k] = dv.x64[(size_t)start * D + k]; q.r[k] = q.x0[k]; }
        u64 rs = mix64(((u64)(u32)start << 8) ^ (u64)attempt ^ 0x51ed270b1f2cULL);
        bool ok = true;
        for (int step = 0; step < D && ok; ++step) {
            double V[D + 1][D];
            unsigned vmask_ = 0;
#pragma unroll
            for (int i = 0; i < D + 1; ++i) {
#pragma unroll
                for (int k = 0; k < D; ++k) V[i][k] = 0.0;
                if (i >= 1 && i < cnt) {
#pragma unroll
                    for (int k = 0; k < D; ++k) V[i][k] = dv.x64[(size_t)sig[i] * D + k] - q.x0[k];
                    vmask_ |= 1u << i;
                }
            }
            double v[D];
#pragma unroll
            for (int k = 0; k < D; ++k) v[k] = unit_hash(rs);
            if (!ortho_direction<D>(V, vmask_, v)) { ok = false; break; }
            Best best;
            best.id = -1; best.t = INFINITY; best.t2 = INFINITY; best.tw = 0.0;
            for (int dir = 0; dir < 2 && best.id < 0; ++dir) {
#pragma unroll
                for (int k = 0; k < D; ++k) q.u[k] = dir ? -v[k] : v[k];
                double cm = -INFINITY;
#pragma unroll
                for (int i = 0; i < D + 1; ++i)
                    if (i < cnt) cm = fmax(cm, dotD<D>(q.u, dv.x64 + (size_t)sig[i] * D));
                q.c = cm + fabs(cm) * dv.plane_tol;
                double w[D];
#pragma unroll
                for (int k = 0; k < D; ++k) w[k] = q.x0[k] - q.r[k];
                q.R0sq = dotD<D>(w, w);
                q.a = dotD<D>(q.u, w);
                q.rnorm = sqrt(dotD<D>(q.r, q.r));
#pragma unroll
                for (int i = 0; i < D + 1; ++i) q.excl[i] = (i < cnt) ? sig[i] : -1;
                q.nexcl = cnt;
                best = min_t_query_call<D, T>(dv, tile, q, ls);
            }
            if (best.id < 0) { ok = false; break; }
#pragma unroll
            for (int k = 0; k < D; ++k) q.r[k] += best.t * q.u[k];
            sig[cnt++] = best.id;
        }
        if (!ok) continue;
        for (int i = 0; i < D + 1; ++i) sorted_insert(ssig, i, sig[i]);
#pragma unroll
        for (int k = 0; k < D; ++k) rr[k] = q.r[k];
        return true;
    }
    return false;
}

// Start of the hull walk (search_max + descent_chull, chull.jl:244-255): `start` is the generator with the largest
// coordinate along `axis`, so the axis direction e lies in the recession cone of its (unbounded) cell.  From a first vertex of
// that cell the walk follows, among the edges that keep `start`, the one that climbs fastest along e -- the simplex method
// on the cell -- until an edge has no end: the first facet.
template <int D, class T>
HVB_HD bool hull_seed(const Dev<D>& dv, const HullDev<D>& hd, const T& tile, int start, int axis, u64* q_out, u32* q_count, u32 q_cap, LocalStats& ls) {
    int sig[D + 1];
    double r[D];
    const int lane = tile.lane();
    if (!descent_vertex<D, T>(dv, tile, start, sig, r, ls)) { ls.seed_fail += (lane == 0); return false; }
    for (int step = 0; step < 100000; ++step) {
        // all lanes of the tile walk the same path (the queries are shared); lane 0 writes, the others read its record
        u32 v = 0;
        if (lane == 0) { v = hull_new_vertex<D>(dv, sig, r); mem_fence(); }
        v = tile.shfl(v, 0);
        if (v == 0xffffffffu) return false;
        int bestk = -1;
        double bests = -INFINITY;
        RayQ<D> qb;
        for (int k = 0; k < D + 1; ++k) {
            if (sig[k] == start) continue;
            RayQ<D> q; int s2[D + 1]; u32 vv; int kk;
            if (!ray_setup<D>(dv, ((u64)v << 3) | (u64)k, q, s2, vv, kk)) continue;
            if (q.u[axis] > bests) { bests = q.u[axis]; bestk = k; qb = q; }
        }
        if (bestk < 0) { ls.seed_fail += (lane == 0); return false; }
        const Best best = min_t_query_call<D, T>(dv, tile, qb, ls);
        if (best.id < 0) {
            if (lane == 0) hull_facet_found<D>(dv, hd, sig, v, bestk, qb.u, q_out, q_count, q_cap);
            return true;
        }
        int e[D];
        for (int j = 0; j < D; ++j) e[j] = (j < bestk) ? sig[j] : sig[j + 1];
        for (int j = 0; j < D; ++j) sig[j] = e[j];
        sorted_insert(sig, D, best.id);
        for (int k = 0; k < D; ++k) r[k] = qb.r[k] + best.t * qb.u[k];
    }
    ls.seed_fail += (lane == 0);
    return false;
}

}  // namespace hvb
