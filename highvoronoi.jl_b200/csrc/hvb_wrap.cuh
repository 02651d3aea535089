// hvb_wrap.cuh -- convex hull by gift wrapping with brute-force, full-warp queries.  Device code that also
// compiles as host C++ (tests/hostsim), like hvb_core.cuh.
//
// Replaces ConvexHull(xs) = systematic_chull (chull.jl:241-387: search_max, descent_chull, then a queue of facets whose
// sub-facets are explored by raycast_des3, chull.jl:485-499, which asks the KD tree for the node that a hyperplane rotating
// about the sub-facet meets first, peak_direction kd_tree.jl:369-420).  Same walk -- facets, ridges, one query per open ridge --
// but the query is re-designed for the machine instead of translated:
//
//   * a tree or grid descent for a half-space is hopeless in SIMT (the first version of this file walked the unbounded 2-faces
//     of the Voronoi diagram with min_t_query: 6 queries per facet, 3 500 candidates per query at 6 of 32 threads; hvb_hull.cuh,
//     kept as hvb_convex_hull_via(ctx, 1)).  Here a query is a STREAM over all generators: the FP32 coordinates are staged
//     tile by tile in shared memory by the TMA unit (cp.async.bulk + mbarrier, k_wrap_scan in hvb_kernels.cuh), every lane of
//     a warp owns one query and all lanes read the same staged point (a shared-memory broadcast): 2 d + 5 FP32 instructions
//     per (query, generator), no divergence, no index.
//   * the pivot rule.  A query is (r0, u, e): the supporting hyperplane {u.(x - r0) = 0} holds the kept generators and all
//     others lie strictly behind it (A(x) = -u.(x - r0) > 0); e is the unit vector in that hyperplane, orthogonal to the flat
//     of the kept generators, pointing away from the generator that leaves.  Rotating the hyperplane about the flat towards e,
//     the first generator it meets maximises  c(x) = B(x) / A(x),  B(x) = e.(x - r0)  (the cotangent of the rotation angle).
//     The new outward normal is u B(g) + e A(g): orientation holds by construction.
//   * exactness as in the vertex search: the FP32 values carry an explicit error bound E; a generator is dropped only if
//     B32 - L A32 + (1 + |L|) E < 0 for a lower bound L of the best cotangent so far, every survivor is evaluated in FP64
//     (wrap_verify); a runner-up within 1e-11 (1 + c^2) of the winner is non-general position (a facet with more than d
//     generators) and is reported like everywhere else.
//   * the first facets come from the same query: start at an extreme generator of an axis, u = +-e_k, and rotate about the
//     flat of the generators found so far (0-, 1-, ..., (d-2)-dimensional) in a hashed direction: d - 1 seed steps.  All 2 d
//     axis extremes start a facet at once (the walk is a breadth-first search over the facet graph whose depth sets the
//     number of rounds, and a round of a small hull is pure latency).
//   * facet set and ridge table are those of hvb_hull.cuh (the vertex set and edge table of the search one dimension down):
//     the first facet at a ridge sends a query around it, the second closes it; an entry whose ridge got closed is skipped.
#pragma once
#include "hvb_hull.cuh"

#ifndef HVB_WRAP_TP
#define HVB_WRAP_TP 256        // generators per staged tile
#endif
#ifndef HVB_WRAP_NS
#define HVB_WRAP_NS 4          // tiles in flight
#endif

namespace hvb {

template <int D>
struct WrapQuery {
    double r0[D], u[D], e[D];
    float nuf[D], ef[D];       // -u and e in FP32
    float ur0, ner0;           // u.(r0 - lo) and -e.(r0 - lo) in FP32: A32 = ur0 + sum nuf x, B32 = ner0 + sum ef x
    int excl[D];               // the kept generators and the one that leaves: never candidates
    int nexcl;
    int pivot;                 // position in excl of the generator that leaves; -1: seed step (all kept)
    u32 src;                   // facet the query starts from, 0xffffffff for a seed step
    u32 seed;                  // seed steps: which of the 2 d first facets is growing
};

struct WrapSeed { int ids[8]; int cnt; int pad; double u[8]; };    // the first facet under construction

template <int D>
struct WrapDev {
    WrapQuery<D>* wq; u32 wq_cap;
    double* pc1; double* pc2; int* pid; u32 pcap;      // partial results, index = chunk * nq + query
    u64* q[2]; u32 qcap;                               // entries: facet << 3 | position of the generator that leaves
    u32* qcount;                                       // [2]
    u32* nq;                                           // [2] live queries of a round (entries minus closed ridges)
    WrapSeed* seed;                                    // [2 D]: one first facet per axis extreme
    float E32;                                         // bound on |A32 - A| and |B32 - B|
    double tinyA;                                      // A <= tinyA: the generator lies in the supporting hyperplane
};
const u64 WRAP_SEED_ENTRY = ~0ULL - 15ULL;           // entries >= this: seed step of seed (entry - WRAP_SEED_ENTRY)

// Error bound of the FP32 values: the coordinates are fl32(x - lo) in [0, ext] (|error| <= 2^-24 ext each), -u and e are
// rounded unit vectors, the sums are FMA chains of d terms of magnitude <= sqrt(d) ext: (2 d + d sqrt(d) + 3 sqrt(d)) 2^-24 ext
// bounds |A32 - A| and |B32 - B|; twice a generous version of it is used.
template <int D>
HVB_HD void wrap_tolerances(double ext, float& E32, double& tinyA) {
    const double eps = 5.9604644775390625e-08;
    E32 = (float)(2.0 * eps * ext * (4.0 * D + (double)D * D + 8.0)) * 1.0001f;
    tinyA = 1e-12 * ext;
}

// How the (query, generator) rectangle of a round is cut into blocks of 4 warps: QW warps own 32 queries each, the other
// PW = 4 / QW split the staged tile; the generators are cut into PCH chunks so that about `tb` blocks exist.  A function of
// device-side counts only, evaluated identically by the scan and the commit kernel: the host never needs the count of a round.
struct WrapShape { int QW, PW, ntile, PCH, chunk; };
HVB_HD WrapShape wrap_shape(u32 nq, int n, int tb, u32 pcap) {
    WrapShape s;
    s.QW = nq > 64 ? 4 : (nq > 32 ? 2 : 1);
    s.PW = 4 / s.QW;
    s.ntile = (int)((nq + 32u * s.QW - 1u) / (32u * s.QW));
    if (s.ntile < 1) s.ntile = 1;
    long long pch = tb / s.ntile;
    const long long maxch = (n + HVB_WRAP_TP - 1) / HVB_WRAP_TP;
    const long long lim = (long long)pcap / (long long)(nq > 0 ? nq : 1);
    if (pch > maxch) pch = maxch;
    if (pch > lim) pch = lim;
    if (pch < 1) pch = 1;
    s.PCH = (int)pch;
    s.chunk = (int)((((long long)n + pch - 1) / pch + 3) & ~3LL);
    return s;
}

template <int D>
struct WrapLane {
    float nuf[D], ef[D], ur0, ner0;
    float L, mL;               // lower bound of the best cotangent, margin (1 + |L|) E
    // FP32 bookkeeping as in scan_points (hvb_core.cuh): the candidate with the largest FP32 LOWER bound of the cotangent and
    // at most one rival whose interval overlaps it; everything else is decided in FP32, FP64 runs for these two at the end
    // of the stream (wrap_settle) and for a third overlapping interval.  id < 0: empty
    int cb_id, cr_id;
    float cb_lo, cb_hi, cr_lo, cr_hi;
    double c1, c2;             // best and runner-up among the FP64 evaluations
    int id1;
};

template <int D>
HVB_HD void wrap_lane_init(const WrapQuery<D>& w, WrapLane<D>& s) {
#pragma unroll
    for (int k = 0; k < D; ++k) { s.nuf[k] = w.nuf[k]; s.ef[k] = w.ef[k]; }
    s.ur0 = w.ur0; s.ner0 = w.ner0;
    s.L = -3.0e38f; s.mL = INFINITY;
    s.cb_id = -1; s.cr_id = -1; s.cb_lo = -3.0e38f; s.cb_hi = -3.0e38f; s.cr_lo = 0.f; s.cr_hi = 0.f;
    s.c1 = -INFINITY; s.c2 = -INFINITY; s.id1 = -1;
}

template <int D>
HVB_HD void wrap_raise(WrapLane<D>& s, float lo, float E) {
    if (lo > s.L) { s.L = lo; s.mL = E * (1.0f + fabsf(lo)) * 1.0001f; }
}

template <int D>
HVB_HD void wrap_offer(WrapLane<D>& s, double c, int p, float E) {
    if (c > s.c1 || (c == s.c1 && p < s.id1)) {
        s.c2 = s.c1; s.c1 = c; s.id1 = p;
        const float f = (float)c;
        wrap_raise<D>(s, f - fabsf(f) * 1.2e-7f, E);   // <= c
    } else if (c > s.c2) s.c2 = c;
}

// FP64 evaluation of one generator
template <int D>
HVB_HD void wrap_verify(const Dev<D>& dv, const WrapQuery<D>& w, double tinyA, float E, int p, WrapLane<D>& s, LocalStats& ls) {
#pragma unroll
    for (int i = 0; i < D; ++i)
        if (i < w.nexcl && w.excl[i] == p) return;
    const double* x = dv.x64 + (size_t)p * D;
    double A = 0, B = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const double dx = x[k] - w.r0[k];
        A -= w.u[k] * dx;
        B += w.e[k] * dx;
    }
    ls.cand64++;
    if (!(A > tinyA)) { ls.degenerate++; return; }     // a generator in (or beyond) the supporting hyperplane
    wrap_offer<D>(s, B / A, p, E);
}

// a generator that passed the filter with FP32 values (A, B), A > 2 E: interval of its cotangent, then the bookkeeping
template <int D>
HVB_HD int wrap_survivor(WrapLane<D>& s, int id, float A, float B, float E) {
    const float bh = B + E, bl = B - E, ah = A + E, al = A - E;          // al > E > 0
    float hi = fast_div(bh, (bh >= 0.f) ? al : ah); hi += fabsf(hi) * 4e-7f;
    float lo = fast_div(bl, (bl >= 0.f) ? ah : al); lo -= fabsf(lo) * 4e-7f;
    int to_verify = -1;
    if (lo > s.cb_lo) {
        // new FP32 best; the old one stays as the rival if its interval still overlaps
        if (s.cb_id >= 0 && s.cb_hi >= lo) {
            if (s.cr_id >= 0 && s.cr_hi >= lo) to_verify = s.cr_id;      // no room: settle the displaced rival now
            s.cr_id = s.cb_id; s.cr_lo = s.cb_lo; s.cr_hi = s.cb_hi;
        } else if (s.cr_id >= 0 && s.cr_hi < lo) s.cr_id = -1;
        s.cb_id = id; s.cb_lo = lo; s.cb_hi = hi;
        wrap_raise<D>(s, lo, E);
    } else if (s.cr_id < 0) { s.cr_id = id; s.cr_lo = lo; s.cr_hi = hi; }
    else to_verify = id;
    return to_verify;
}

// end of a stream: the FP64 evaluation of the FP32 winner and of a surviving rival
template <int D>
HVB_HD void wrap_settle(const Dev<D>& dv, const WrapQuery<D>& w, double tinyA, float E, WrapLane<D>& s, LocalStats& ls) {
    if (s.cb_id >= 0) wrap_verify<D>(dv, w, tinyA, E, s.cb_id, s, ls);
    if (s.cr_id >= 0 && s.cr_hi >= s.L) wrap_verify<D>(dv, w, tinyA, E, s.cr_id, s, ls);
    s.cb_id = -1; s.cr_id = -1; s.cb_lo = -3.0e38f; s.cb_hi = -3.0e38f;
}

template <int D>
HVB_HD void wrap_load(const float* p, float (&x)[D]) {
#if defined(__CUDA_ARCH__)
    float t[8];
    if (D == 2) {
        const float2 v = *reinterpret_cast<const float2*>(p);
        t[0] = v.x; t[1] = v.y;
    } else if (D <= 4) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    } else {
        const float4 v = *reinterpret_cast<const float4*>(p);
        const float4 w = *reinterpret_cast<const float4*>(p + 4);
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w; t[4] = w.x; t[5] = w.y; t[6] = w.z; t[7] = w.w;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = t[k];
#else
    for (int k = 0; k < D; ++k) x[k] = p[k];
#endif
}

// the stream: `count` generators at `pts` (FP32, stride X32<D>::STRIDE; readable up to the next multiple of four), ids from
// `first_id`.  On the device `pts` is a staged tile in shared memory and all lanes of a warp read the same address.
// wrap_settle() must follow the last call.
template <int D>
HVB_HD void wrap_scan(const Dev<D>& dv, const WrapQuery<D>& w, const float* pts, int first_id, int count, double tinyA, float E,
                      WrapLane<D>& s, LocalStats& ls) {
    const int S = X32<D>::STRIDE;
    const float E2 = 2.f * E;
    // Branch-free pre-pass over every fourth generator: the largest FP32 lower bound of a cotangent among them raises L
    // before the stream starts.  The generators arrive in grid order, i.e. spatially coherent: along a grid row the
    // cotangent of a query can rise monotonically, and then every generator is a new best that takes the (serial, branchy)
    // survivor path -- measured 90 us for the 1700 generators of a chunk in d = 2.  After the pre-pass only what beats the
    // best of the sample survives: about four generators whatever the order.
    {
        float best = s.L;
        for (int i = 0; i < count; i += 16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ii = (i + 4 * j < count) ? i + 4 * j : 0;
                float x[D];
                wrap_load<D>(pts + (size_t)ii * S, x);
                float A = s.ur0, B = s.ner0;
#pragma unroll
                for (int k = 0; k < D; ++k) { A = fmaf(s.nuf[k], x[k], A); B = fmaf(s.ef[k], x[k], B); }
                const float bl = B - E;
                float lo = fast_div(bl, (bl >= 0.f) ? A + E : A - E);
                lo -= fabsf(lo) * 4e-7f;
                best = (A > E2 && lo > best) ? lo : best;
            }
        }
        wrap_raise<D>(s, best, E);
    }
    for (int i = 0; i < count; i += 4) {
        unsigned pm = 0;
        float As[4], Bs[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x[D];
            wrap_load<D>(pts + (size_t)(i + j) * S, x);
            float A = s.ur0, B = s.ner0;
#pragma unroll
            for (int k = 0; k < D; ++k) { A = fmaf(s.nuf[k], x[k], A); B = fmaf(s.ef[k], x[k], B); }
            As[j] = A; Bs[j] = B;
            const bool pass = (i + j < count) && (!(A > E2) || fmaf(-s.L, A, B) + s.mL >= 0.f);
            pm |= pass ? (1u << j) : 0u;
        }
        while (pm) {
            const int j = lowest_bit(pm);
            pm &= pm - 1u;
            float A = As[0], B = Bs[0];
#pragma unroll
            for (int b = 1; b < 4; ++b) { A = (j == b) ? As[b] : A; B = (j == b) ? Bs[b] : B; }
            int to_verify = first_id + i + j;
            if (A > E2) {
                if (!(fmaf(-s.L, A, B) + s.mL >= 0.f)) continue;         // the bound may have risen within the chunk
                to_verify = wrap_survivor<D>(s, first_id + i + j, A, B, E);
            }
            if (to_verify >= 0) wrap_verify<D>(dv, w, tinyA, E, to_verify, s, ls);
        }
    }
}

// merge of partial results (disjoint generator sets)
HVB_HD void wrap_merge(double& c1, int& g, double& c2, double oc1, int og, double oc2) {
    if (og >= 0 && (g < 0 || oc1 > c1 || (oc1 == c1 && og < g))) {
        c2 = fmax(fmax(c2, oc2), (g >= 0) ? c1 : -INFINITY);
        c1 = oc1; g = og;
    } else {
        c2 = fmax(c2, fmax((og >= 0) ? oc1 : -INFINITY, oc2));
    }
}

// is the ridge (E minus position j) closed, i.e. are both its facets known?
template <int D>
HVB_HD bool ridge_closed(const HullDev<D>& hd, const int* E, int j) {
    const u64 h = hash_facet<D>(E, j);
    const u64 probe = edge_slot(h, 0, 0);
    u64 slot = h & hd.rmask;
    for (;;) {
        const u64 s = ld_cg(hd.rtab + slot);
        if (s == 0) return false;
        if (((s ^ probe) & EDGE_FPMASK) == 0) {
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
            if (eq) return (s & EDGE_CLOSED) != 0;
        }
        slot = (slot + 1) & hd.rmask;
    }
}

// entry -> query.  false: nothing to do (ridge closed meanwhile, dead facet record)
template <int D>
HVB_HD bool wrap_prepare(const Dev<D>& dv, const HullDev<D>& hd, const WrapDev<D>& wd, u64 entry, WrapQuery<D>& w, LocalStats& ls) {
    double V[D + 1][D];
#pragma unroll
    for (int i = 0; i < D + 1; ++i)
#pragma unroll
        for (int k = 0; k < D; ++k) V[i][k] = 0.0;
    unsigned mask = 1u;
    double v[D];
    if (entry >= WRAP_SEED_ENTRY) {
        const u32 si = (u32)(entry - WRAP_SEED_ENTRY);
        const WrapSeed& sd = wd.seed[si];
        const int cnt = sd.cnt;
        w.nexcl = cnt; w.pivot = -1; w.src = 0xffffffffu; w.seed = si;
#pragma unroll
        for (int i = 0; i < D; ++i) w.excl[i] = (i < cnt) ? sd.ids[i] : -1;
#pragma unroll
        for (int k = 0; k < D; ++k) { w.r0[k] = dv.x64[(size_t)sd.ids[0] * D + k]; w.u[k] = sd.u[k]; }
        bool ok = false;
        u64 rs = mix64(0x5eed0a11ULL + 977ULL * (u64)cnt + 7919ULL * (u64)si);
        for (int attempt = 0; attempt < 8 && !ok; ++attempt) {
#pragma unroll
            for (int k = 0; k < D; ++k) V[0][k] = w.u[k];
            mask = 1u;
#pragma unroll
            for (int i = 1; i < D; ++i) {
                if (i < cnt) {
#pragma unroll
                    for (int k = 0; k < D; ++k) V[i][k] = dv.x64[(size_t)sd.ids[i] * D + k] - w.r0[k];
                    mask |= 1u << i;
                }
            }
            double nv = 0;
#pragma unroll
            for (int k = 0; k < D; ++k) { v[k] = unit_hash(rs); nv += v[k] * v[k]; }
            double before[D];
#pragma unroll
            for (int k = 0; k < D; ++k) before[k] = v[k];
            ok = ortho_direction<D>(V, mask, v) && fabs(dotD<D>(before, v)) > 1e-3 * sqrt(nv);      // not a direction almost inside the flat
        }
        if (!ok) { ls.seed_fail++; return false; }
    } else {
        const u32 f = (u32)(entry >> 3);
        const int j = (int)(entry & 7);
        int F[D];
#pragma unroll
        for (int i = 0; i < D; ++i) F[i] = ld_cg(hd.fsig + (size_t)f * D + i);
        if (F[0] < 0) return false;
        if (ridge_closed<D>(hd, F, j)) { ls.closed_skips++; return false; }
        const int b = (j == 0) ? 1 : 0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            w.r0[k] = dv.x64[(size_t)F[b] * D + k];
            w.u[k] = ld_cg(hd.fu + (size_t)f * D + k);
            V[0][k] = w.u[k];
        }
        int row = 1;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            if (i == j || i == b) continue;
#pragma unroll
            for (int r = 1; r < D; ++r)
                if (r == row) {
#pragma unroll
                    for (int k = 0; k < D; ++k) V[r][k] = dv.x64[(size_t)F[i] * D + k] - w.r0[k];
                }
            mask |= 1u << row;
            ++row;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = w.r0[k] - dv.x64[(size_t)F[j] * D + k];
        if (!ortho_direction<D>(V, mask, v)) { ls.degenerate++; return false; }
#pragma unroll
        for (int i = 0; i < D; ++i) w.excl[i] = F[i];
        w.nexcl = D; w.pivot = j; w.src = f; w.seed = 0;
    }
    double ur0 = 0, er0 = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        w.e[k] = v[k];
        w.nuf[k] = (float)(-w.u[k]);
        w.ef[k] = (float)v[k];
        ur0 += w.u[k] * (w.r0[k] - dv.lo[k]);
        er0 += v[k] * (w.r0[k] - dv.lo[k]);
    }
    w.ur0 = (float)ur0; w.ner0 = (float)(-er0);
    return true;
}

// the answer of a query: a new facet (or the next seed step)
template <int D>
HVB_HD void wrap_commit(const Dev<D>& dv, const HullDev<D>& hd, const WrapDev<D>& wd, const WrapQuery<D>& w, double c1, int g, double c2,
                        int nxt, LocalStats& ls) {
    if (g < 0) { ls.seed_fail++; return; }
    if (!(c1 - c2 > 1e-11 * (1.0 + c1 * c1))) ls.degenerate++;            // two generators met at the same angle
    const double* xg = dv.x64 + (size_t)g * D;
    double A = 0, B = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const double dx = xg[k] - w.r0[k];
        A -= w.u[k] * dx;
        B += w.e[k] * dx;
    }
    double v[D];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = w.u[k] * B + w.e[k] * A;
    int ids[D];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < D; ++i)
        if (i < w.nexcl && i != w.pivot) {
#pragma unroll
            for (int r = 0; r < D; ++r) if (r == cnt) ids[r] = w.excl[i];
            ++cnt;
        }
#pragma unroll
    for (int r = 0; r < D; ++r) if (r == cnt) ids[r] = g;
    ++cnt;
    if (cnt < D) {
        // the first facet is still growing: rotate about the larger flat next
        const double inv = inv_sqrt(dotD<D>(v, v));
        WrapSeed& sd = wd.seed[w.seed];
#pragma unroll
        for (int i = 0; i < D; ++i) if (i < cnt) sd.ids[i] = ids[i];
        sd.cnt = cnt;
#pragma unroll
        for (int k = 0; k < D; ++k) sd.u[k] = v[k] * inv;
        mem_fence();
        hull_push(wd.q[nxt], wd.qcount + nxt, wd.qcap, &dv.ctr->flags, WRAP_SEED_ENTRY + (u64)w.seed);
        return;
    }
    if (w.pivot < 0) {
        // seed ids arrive in the order they were found
        for (int a = 1; a < D; ++a) { const int key = ids[a]; int b = a - 1; while (b >= 0 && ids[b] > key) { ids[b + 1] = ids[b]; --b; } ids[b + 1] = key; }
    } else {
        // the kept generators are sorted already: insert the winner
        int pos = 0;
#pragma unroll
        for (int i = 0; i < D - 1; ++i) pos += (ids[i] < g) ? 1 : 0;
        int s2[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const int lo_ = (i < D - 1) ? ids[i] : 0, hi_ = (i > 0) ? ids[i - 1] : 0;
            s2[i] = (i < pos) ? lo_ : ((i == pos) ? g : hi_);
        }
#pragma unroll
        for (int i = 0; i < D; ++i) ids[i] = s2[i];
    }
    double V[D + 1][D];
    unsigned mask = 0;
#pragma unroll
    for (int i = 0; i < D + 1; ++i) {
#pragma unroll
        for (int k = 0; k < D; ++k) V[i][k] = 0.0;
        if (i >= 1 && i < D) {
#pragma unroll
            for (int k = 0; k < D; ++k) V[i][k] = dv.x64[(size_t)ids[i] * D + k] - dv.x64[(size_t)ids[0] * D + k];
            mask |= 1u << i;
        }
    }
    if (!ortho_direction<D>(V, mask, v)) { ls.degenerate++; return; }
    const u32 f = facet_insert<D>(dv, hd, ids, 0u, 0, v);
    if (f == 0xffffffffu) { ls.dup_hits++; return; }
#pragma unroll
    for (int j = 0; j < D; ++j)
        if (ridge_register<D>(hd, ids, f, j)) hull_push(wd.q[nxt], wd.qcount + nxt, wd.qcap, &dv.ctr->flags, ((u64)f << 3) | (u64)j);
}

// canonical output of a facet, from its generators alone in ascending CALLER order (like canonical_vertex: independent of
// the path that found it): outward unit normal (orientation from the stored one) and the circumcentre of the d generators
// inside the facet's hyperplane (the point the reference reports, chull.jl:224-232)
template <int D>
HVB_HD bool wrap_facet_geometry(const double (&P)[D][D], const double* fu, double (&nrm)[D], double (&cen)[D]) {
    double V[D + 1][D];
    unsigned mask = 0;
#pragma unroll
    for (int i = 0; i < D + 1; ++i) {
#pragma unroll
        for (int k = 0; k < D; ++k) V[i][k] = 0.0;
        if (i >= 1 && i < D) {
#pragma unroll
            for (int k = 0; k < D; ++k) V[i][k] = P[i][k] - P[0][k];
            mask |= 1u << i;
        }
    }
    // circumcentre: c = p0 + sum lambda_i w_i,  (w_i . w_j) lambda = |w_i|^2 / 2
    double G[D][D + 1];
    const int m = D - 1;
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < m; ++j) G[i][j] = dotD<D>(V[i + 1], V[j + 1]);
        G[i][m] = 0.5 * G[i][i];
    }
    for (int c = 0; c < m; ++c) {
        int piv = c;
        for (int r = c + 1; r < m; ++r) if (fabs(G[r][c]) > fabs(G[piv][c])) piv = r;
        if (piv != c) for (int k = 0; k <= m; ++k) { const double t = G[c][k]; G[c][k] = G[piv][k]; G[piv][k] = t; }
        if (!(fabs(G[c][c]) > 0)) return false;
        for (int r = c + 1; r < m; ++r) {
            const double fct = G[r][c] / G[c][c];
            for (int k = c; k <= m; ++k) G[r][k] -= fct * G[c][k];
        }
    }
    double lam[D];
    for (int c = m - 1; c >= 0; --c) {
        double s = G[c][m];
        for (int k = c + 1; k < m; ++k) s -= G[c][k] * lam[k];
        lam[c] = s / G[c][c];
    }
    for (int k = 0; k < D; ++k) {
        double s = P[0][k];
        for (int i = 0; i < m; ++i) s += lam[i] * V[i + 1][k];
        cen[k] = s;
    }
    // normal: the axis the stored normal is largest along, projected off the facet's directions
    int kmax = 0;
    for (int k = 1; k < D; ++k) if (fabs(fu[k]) > fabs(fu[kmax])) kmax = k;
    for (int k = 0; k < D; ++k) nrm[k] = (k == kmax) ? ((fu[kmax] < 0) ? -1.0 : 1.0) : 0.0;
    return ortho_direction<D>(V, mask, nrm);
}

}  // namespace hvb
