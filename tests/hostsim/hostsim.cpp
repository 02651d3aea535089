// hostsim.cpp -- DEBUGGING HARNESS (test infrastructure, never shipped, never a fallback).
// Compiles the device algorithm of highvoronoi.jl_b200/csrc/hvb_core.cuh as plain host C++ with a one-lane
// tile and drives it sequentially, so that its logic (grid traversal, FP32 filter bound, probe stages,
// vertex/edge tables, descent) can be checked against the oracle in a container that has no GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
// per-ray trace of the query (kind: 0 ray, 1 row test, 2 row scanned with val points, 3 stage with val rows, 4 FP32 survivor)
static std::vector<int>* g_trace = nullptr;
// Iter subset / slab of a multi-GPU run: cells (caller order) the next run explores; empty = all (hostsim_set_active)
static std::vector<unsigned char> g_active;
static bool g_want_trace = false;
#define HVB_TRACE_EVENT(kind, val) do { if (g_trace) { g_trace->push_back(kind); g_trace->push_back((int)(val)); } } while (0)
#include "../../highvoronoi.jl_b200/csrc/hvb_host.hpp"
#include "../../highvoronoi.jl_b200/csrc/hvb_geometry.cuh"
#include "../../highvoronoi.jl_b200/csrc/hvb_hull.cuh"
#include "../../highvoronoi.jl_b200/csrc/hvb_wrap.cuh"
#include "../../highvoronoi.jl_b200/csrc/hvb_nongeneral.hpp"

using namespace hvb;

struct SimResult {
    int d; int64_t nv, nr;
    std::vector<int64_t> sig; std::vector<double> r;
    std::vector<int64_t> ray_edge;
    Counters ctr;
    int rounds;
    std::vector<int> trace;
};

template <int D>
static SimResult* run(int64_t n, const double* xs, int P, const double* base, const double* normal,
                      int ppc, double probe_scale, int fp32, int seed_stride,
                      const double* xs_canon = nullptr, double t_min = 1e-12, double flat_tol = 0.0) {
    // xs_canon / t_min / flat_tol: the search of a PERTURBED cloud (resolve_degenerate): coordinates are solved from the
    // caller's generators xs_canon, the smallest ray parameter is t_min, simplices flat in xs_canon are dropped
    Dev<D> dv;
    memset(&dv, 0, sizeof(dv));
    dv.n = (int)n;
    double blo[D], bhi[D];
    for (int k = 0; k < D; ++k) { blo[k] = 1e300; bhi[k] = -1e300; }
    for (int64_t i = 0; i < n; ++i)
        for (int k = 0; k < D; ++k) { blo[k] = std::min(blo[k], xs[i * D + k]); bhi[k] = std::max(bhi[k], xs[i * D + k]); }
    // HOSTSIM_CELLS_MULT: experiments with finer grids than one cell per `ppc` points (clustered clouds)
    const double cells_mult = getenv("HOSTSIM_CELLS_MULT") ? atof(getenv("HOSTSIM_CELLS_MULT")) : 1.0;
    int64_t ncell = setup_grid<D>(dv, blo, bhi, (int64_t)(n * cells_mult), ppc > 0 ? ppc : default_points_per_cell(D));
    std::vector<int> cell(n), cstart(ncell + 1, 0), perm(n);
    for (int64_t i = 0; i < n; ++i) { cell[i] = cell_index<D>(dv, xs + i * D); cstart[cell[i] + 1]++; }
    for (int64_t c = 0; c < ncell; ++c) cstart[c + 1] += cstart[c];
    { std::vector<int> cur(cstart.begin(), cstart.end() - 1); for (int64_t i = 0; i < n; ++i) perm[cur[cell[i]]++] = (int)i; }
    std::vector<double> x64(n * D); std::vector<float> x32(n * X32<D>::STRIDE, 0.f);
    for (int64_t i = 0; i < n; ++i)
        for (int k = 0; k < D; ++k) { x64[i * D + k] = xs[(int64_t)perm[i] * D + k]; x32[i * X32<D>::STRIDE + k] = (float)(x64[i * D + k] - dv.lo[k]); }
    PlaneSet ps; memset(&ps, 0, sizeof(ps)); ps.P = P;
    for (int p = 0; p < P; ++p) {
        double nr = 0; for (int k = 0; k < D; ++k) nr += normal[p * D + k] * normal[p * D + k];
        nr = sqrt(nr); double off = 0;
        for (int k = 0; k < D; ++k) { ps.normal[p * 6 + k] = normal[p * D + k] / nr; off += ps.normal[p * 6 + k] * base[p * D + k]; }
        ps.off[p] = off;
    }
    std::vector<unsigned char> active(n, 1), hasv(n, 0);
    if ((int64_t)g_active.size() == n) for (int64_t i = 0; i < n; ++i) active[i] = g_active[perm[i]];   // grid order
    std::vector<double> xcan;
    if (xs_canon) { xcan.resize(n * D); for (int64_t i = 0; i < n; ++i) for (int k = 0; k < D; ++k) xcan[i * D + k] = xs_canon[(int64_t)perm[i] * D + k]; }
    dv.cell_start = cstart.data(); dv.x32 = x32.data(); dv.x64 = x64.data(); dv.xcan = xs_canon ? xcan.data() : x64.data(); dv.planes = &ps; dv.active = active.data();
    dv.plane_tol = 1e-12; dv.t_min = t_min; dv.probe_scale = probe_scale > 1.0 ? probe_scale : 1.3; dv.probe_growth = 2.0; dv.fp32_filter = fp32;
    int64_t vcap = estimate_vertices(D, n, P) * 2;
    std::vector<int> vsig(vcap * (D + 1)); std::vector<double> vr(vcap * D);
    u32 vcount = 0;
    u64 vts = next_pow2(2 * vcap), ets = next_pow2((u64)(vcap * (D + 1)));
    std::vector<u64> vtab(vts, 0), etab(ets, 0);
    u32 rcount = 0; u32 rcap = (u32)vcap;
    std::vector<u32> ritem(rcap); std::vector<double> ru((size_t)rcap * D);
    Counters ctr; memset(&ctr, 0, sizeof(ctr));
    dv.vsig = vsig.data(); dv.vr = vr.data(); dv.vcount = &vcount; dv.vcap = (u32)vcap; dv.vtab = vtab.data(); dv.vmask = vts - 1;
    dv.etab = etab.data(); dv.emask = ets - 1; dv.has_vertex = hasv.data();
    dv.ray_item = ritem.data(); dv.ray_u = ru.data(); dv.ray_count = &rcount; dv.ray_cap = rcap; dv.ctr = &ctr;
    u32 qcap = (u32)std::min<u64>(ets, 0xfffffff0u);
    std::vector<u64> qa(qcap), qb(qcap);
    u32 na = 0, nb = 0;
    TileHost tile; LocalStats ls; memset(&ls, 0, sizeof(ls));
    if (seed_stride <= 0) seed_stride = 16;
    std::vector<int> trace_buf;
    if (g_want_trace) g_trace = &trace_buf;
    int64_t watch[D + 1]; int watch_n = 0;
    const bool watch_all = getenv("HOSTSIM_WATCH") && !strcmp(getenv("HOSTSIM_WATCH"), "all");
    if (const char* w = watch_all ? nullptr : getenv("HOSTSIM_WATCH")) {
        for (const char* c = w; *c && watch_n <= D; ) { watch[watch_n++] = strtoll(c, (char**)&c, 10); if (*c == ',') ++c; }
        std::sort(watch, watch + watch_n);
        g_trace = &trace_buf;
    }
    if (watch_all) g_trace = &trace_buf;
    for (int64_t i = 0; i < n; i += seed_stride) {
        const u32 v_before = vcount;
        if (!active[i]) continue;                          // k_seed: descents start from cells this context explores
        seed_item<D, TileHost>(dv, tile, (int)i, qa.data(), &na, qcap, ls);
        if (watch_all && vcount > v_before) {
            fprintf(stderr, "[seed] start %lld -> (", (long long)perm[i] + 1);
            for (int k = 0; k <= D; ++k) { int id = vsig[(size_t)v_before * (D + 1) + k]; fprintf(stderr, "%lld ", (long long)((id < n ? perm[id] : id) + 1)); }
            fprintf(stderr, ")\n");
        }
    }
    int rounds = 0;
    for (;;) {
        while (na > 0) {
            nb = 0;
            for (u32 i = 0; i < na; ++i) {
                if (etab[(u32)(qa[i] >> 32)] & EDGE_CLOSED) { ls.closed_skips++; continue; }   // as in k_expand
                const u32 v_before = vcount;
                const size_t tr_before = trace_buf.size();
                expand_item<D, TileHost>(dv, tile, qa[i], qb.data(), &nb, qcap, ls);
                if ((watch_n == D + 1 || watch_all) && vcount > v_before) {
                    // HOSTSIM_WATCH=id,id,...: report the walk that produced this vertex (caller ids, 1-based)
                    int64_t s[D + 1];
                    for (int k = 0; k <= D; ++k) { int id = vsig[(size_t)v_before * (D + 1) + k]; s[k] = (id < n ? perm[id] : id) + 1; }
                    std::sort(s, s + D + 1);
                    bool same = true; for (int k = 0; k <= D; ++k) same &= (s[k] == watch[k]);
                    if (same || watch_all) {
                        const u32 vo = (u32)(qa[i] & 0xffffffffu) >> 3; const int kd = (int)(qa[i] & 7);
                        fprintf(stderr, "[watch] (");
                        for (int k = 0; k <= D; ++k) fprintf(stderr, "%lld ", (long long)s[k]);
                        fprintf(stderr, ") found from vertex (");
                        for (int k = 0; k <= D; ++k) { int id = vsig[(size_t)vo * (D + 1) + k]; fprintf(stderr, "%lld%s", (long long)((id < n ? perm[id] : id) + 1), k == kd ? "* " : " "); }
                        fprintf(stderr, ") r = (");
                        for (int k = 0; k < D; ++k) fprintf(stderr, "%.17g ", vr[(size_t)vo * D + k]);
                        fprintf(stderr, ")  grid g = (");
                        for (int k = 0; k < D; ++k) fprintf(stderr, "%d ", dv.g[k]);
                        fprintf(stderr, ") h = (");
                        for (int k = 0; k < D; ++k) fprintf(stderr, "%.4g ", dv.h[k]);
                        fprintf(stderr, ")\n[watch] trace:");
                        for (size_t t = tr_before; t + 1 < trace_buf.size(); t += 2) fprintf(stderr, " %d:%d", trace_buf[t], trace_buf[t + 1]);
                        fprintf(stderr, "\n");
                    }
                }
            }
            qa.swap(qb); na = nb; ++rounds;
        }
        // cells without any vertex get their own descent (sysvoronoi.jl:416-429)
        for (int64_t i = 0; i < n; ++i) if (active[i] && !hasv[i]) seed_item<D, TileHost>(dv, tile, (int)i, qa.data(), &na, qcap, ls);
        if (na == 0) break;
    }
    g_trace = nullptr;
    SimResult* R = new SimResult(); R->d = D; R->rounds = rounds;
    R->trace.swap(trace_buf);
    ctr.raycasts = ls.raycasts; ctr.dup_hits = ls.dup_hits; ctr.closed_skips = ls.closed_skips; ctr.cand32 = ls.cand32; ctr.cand64 = ls.cand64;
    ctr.rows = ls.rows; ctr.stages = ls.stages; ctr.seeds = ls.seeds; ctr.degenerate = ls.degenerate; ctr.seed_fail = ls.seed_fail; ctr.dead = ls.dead;
    R->ctr = ctr;
    // finalize: caller ids, canonical order, canonical coordinates
    struct Row { int64_t sig[7]; double r[6]; };
    std::vector<Row> rows;
    for (u32 v = 0; v < vcount; ++v) {
        const int* s = &vsig[(size_t)v * (D + 1)];
        if (s[0] < 0) continue;
        std::pair<int64_t, int> o[D + 1];
        for (int k = 0; k <= D; ++k) o[k] = std::make_pair(s[k] < n ? (int64_t)perm[s[k]] : (int64_t)s[k], s[k]);
        std::sort(o, o + D + 1);
        int cs[D + 1]; Row row;
        for (int k = 0; k <= D; ++k) { cs[k] = o[k].second; row.sig[k] = o[k].first + 1; }
        double flat = 1.0;
        canonical_vertex<D>(dv, cs, row.r, &flat);
        if (flat_tol > 0 && !(flat > flat_tol)) continue;            // a sliver of the perturbed triangulation (k_final_rows)
        rows.push_back(row);
    }
    std::sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return std::lexicographical_compare(a.sig, a.sig + D + 1, b.sig, b.sig + D + 1); });
    R->nv = (int64_t)rows.size();
    for (auto& row : rows) { for (int k = 0; k <= D; ++k) R->sig.push_back(row.sig[k]); for (int k = 0; k < D; ++k) R->r.push_back(row.r[k]); }
    R->nr = rcount;
    for (u32 i = 0; i < rcount; ++i) {
        u32 v = ritem[i] >> 3; int kd = ritem[i] & 7;
        std::vector<int64_t> e;
        for (int k = 0; k <= D; ++k) if (k != kd) { int id = vsig[(size_t)v * (D + 1) + k]; e.push_back((id < n ? perm[id] : id) + 1); }
        std::sort(e.begin(), e.end());
        R->ray_edge.insert(R->ray_edge.end(), e.begin(), e.end());
    }
    return R;
}

// Non-general position resolved as Ctx::resolve_degenerate does it (hvb_ctx.cuh): perturb, search, coordinates from the caller's
// generators, slivers dropped, rows merged, neighbour lists by full interfaces -- with the library's own host functions
struct ResolveResult { std::vector<int64_t> off, ids, nb_off, nb_ids; std::vector<double> r; int64_t maxlen, ndeg, nsimplicial; };
template <int D>
static ResolveResult* run_resolve(int64_t n, const double* xs, int P, const double* base, const double* normal) {
    double ext = 0;
    for (int k = 0; k < D; ++k) { double lo = 1e300, hi = -1e300; for (int64_t i = 0; i < n; ++i) { lo = std::min(lo, xs[i * D + k]); hi = std::max(hi, xs[i * D + k]); } ext = std::max(ext, hi - lo); }
    std::vector<double> xp(n * D);
    for (int64_t i = 0; i < n * D; ++i) xp[i] = xs[i] + HVB_PERTURB_REL * ext * perturb_unit((u64)i);
    SimResult* S = run<D>(n, xp.data(), P, base, normal, 0, 0.0, 1, 0, xs, HVB_TMIN_REL * ext, HVB_FLAT_TOL);
    ResolveResult* R = new ResolveResult();
    R->nsimplicial = S->nv;
    merge_rows(D, S->nv, S->sig.data(), S->r.data(), HVB_MERGE_REL * ext, true, R->off, R->ids, R->r, R->maxlen, R->ndeg);
    merged_neighbors(D, n, (int64_t)R->off.size() - 1, R->off.data(), R->ids.data(), R->r.data(), 0, nullptr, nullptr, R->nb_off, R->nb_ids);
    delete S;
    return R;
}

// cell volumes from vertex rows (vertex_flag_sum, hvb_geometry.cuh): sig [nv][dim+1] caller ids, 1-based, plane p = n + p
template <int D>
static void volumes(int64_t n, const double* xs, int P, const double* base, const double* normal, int64_t nv, const int64_t* sig, double* vol) {
    PlaneSet ps; memset(&ps, 0, sizeof(ps)); ps.P = P;
    for (int p = 0; p < P; ++p) {
        double nr = 0; for (int k = 0; k < D; ++k) nr += normal[p * D + k] * normal[p * D + k];
        nr = sqrt(nr); double off = 0;
        for (int k = 0; k < D; ++k) { ps.normal[p * 6 + k] = normal[p * D + k] / nr; off += ps.normal[p * 6 + k] * base[p * D + k]; }
        ps.off[p] = off;
    }
    double fact = 1; for (int k = 2; k <= D; ++k) fact *= k;
    for (int64_t i = 0; i < n; ++i) vol[i] = 0;
    for (int64_t v = 0; v < nv; ++v) {
        long long s[D + 1];
        for (int k = 0; k <= D; ++k) s[k] = sig[v * (D + 1) + k];
        for (int k = 0; k <= D; ++k) if (s[k] <= n) vol[s[k] - 1] += vertex_flag_sum<D>(xs, n, &ps, s, k) / fact;
    }
}
// integrals of 1, x_a, x_a x_b over the cells (global coordinates) from vertex rows: out[n][1 + D + D(D+1)/2]
template <int D>
static void moments(int64_t n, const double* xs, int P, const double* base, const double* normal, int64_t nv, const int64_t* sig, double* out) {
    PlaneSet ps; memset(&ps, 0, sizeof(ps)); ps.P = P;
    for (int p = 0; p < P; ++p) {
        double nn = 0; for (int k = 0; k < D; ++k) nn += normal[p * D + k] * normal[p * D + k];
        nn = sqrt(nn); double off = 0;
        for (int k = 0; k < D; ++k) { ps.normal[p * 6 + k] = normal[p * D + k] / nn; off += ps.normal[p * 6 + k] * base[p * D + k]; }
        ps.off[p] = off;
    }
    const int NM = 1 + D + D * (D + 1) / 2;
    double fact = 1.0; for (int k = 2; k <= D; ++k) fact *= k;
    std::vector<double> loc((size_t)n * NM, 0.0);
    for (int64_t v = 0; v < nv; ++v) {
        const int64_t* s = sig + v * (D + 1);
        long long ss[D + 1]; for (int k = 0; k < D + 1; ++k) ss[k] = s[k];
        for (int k = 0; k < D + 1; ++k) {
            if (ss[k] > n) continue;
            double m[NM];
            vertex_flag_moments<D>(xs, n, &ps, ss, k, m);
            for (int a = 0; a < NM; ++a) loc[(size_t)(ss[k] - 1) * NM + a] += m[a] / fact;
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        const double* l = &loc[(size_t)i * NM]; const double* x = xs + i * D; double* o = out + i * NM;
        o[0] = l[0];
        for (int a = 0; a < D; ++a) o[1 + a] = x[a] * l[0] + l[1 + a];
        int q = 0;
        for (int a = 0; a < D; ++a) for (int b = a; b < D; ++b, ++q) o[1 + D + q] = x[a] * x[b] * l[0] + x[a] * l[1 + b] + x[b] * l[1 + a] + l[1 + D + q];
    }
}
// interface areas aligned with the CSR neighbour lists (off[n+1], ids ascending per cell)
template <int D>
static void areas(int64_t n, const double* xs, int P, const double* base, const double* normal, int64_t nv, const int64_t* sig,
                  const int64_t* off, const int64_t* ids, double* area) {
    PlaneSet ps; memset(&ps, 0, sizeof(ps)); ps.P = P;
    for (int p = 0; p < P; ++p) {
        double nr = 0; for (int k = 0; k < D; ++k) nr += normal[p * D + k] * normal[p * D + k];
        nr = sqrt(nr); double o = 0;
        for (int k = 0; k < D; ++k) { ps.normal[p * 6 + k] = normal[p * D + k] / nr; o += ps.normal[p * 6 + k] * base[p * D + k]; }
        ps.off[p] = o;
    }
    double fact = 1; for (int k = 2; k <= D - 1; ++k) fact *= k;
    for (int64_t i = 0; i < off[n]; ++i) area[i] = 0;
    for (int64_t v = 0; v < nv; ++v) {
        long long s[D + 1];
        for (int k = 0; k <= D; ++k) s[k] = sig[v * (D + 1) + k];
        for (int k = 0; k <= D; ++k) {
            if (s[k] > n) continue;
            for (int q = 0; q <= D; ++q) {
                if (q == k) continue;
                const int64_t* a = ids + off[s[k] - 1]; const int64_t* b = ids + off[s[k]];
                const int64_t* it = std::lower_bound(a, b, (int64_t)s[q]);
                if (it == b || *it != s[q]) continue;
                area[it - ids] += vertex_flag_sum<D>(xs, n, &ps, s, k, q) / fact;
            }
        }
    }
}

// area and first moment (global coordinates) of every interface, aligned with the CSR neighbour lists: out[off[n]][1 + D]
template <int D>
static void area_moments(int64_t n, const double* xs, int P, const double* base, const double* normal, int64_t nv, const int64_t* sig,
                         const int64_t* off, const int64_t* ids, double* out) {
    PlaneSet ps; memset(&ps, 0, sizeof(ps)); ps.P = P;
    for (int p = 0; p < P; ++p) {
        double nr = 0; for (int k = 0; k < D; ++k) nr += normal[p * D + k] * normal[p * D + k];
        nr = sqrt(nr); double o = 0;
        for (int k = 0; k < D; ++k) { ps.normal[p * 6 + k] = normal[p * D + k] / nr; o += ps.normal[p * 6 + k] * base[p * D + k]; }
        ps.off[p] = o;
    }
    const int NM = 1 + D + D * (D + 1) / 2;
    double fact = 1; for (int k = 2; k <= D - 1; ++k) fact *= k;
    for (int64_t i = 0; i < off[n] * (1 + D); ++i) out[i] = 0;
    for (int64_t v = 0; v < nv; ++v) {
        long long s[D + 1];
        for (int k = 0; k <= D; ++k) s[k] = sig[v * (D + 1) + k];
        for (int k = 0; k <= D; ++k) {
            if (s[k] > n) continue;
            for (int q = 0; q <= D; ++q) {
                if (q == k) continue;
                const int64_t* a = ids + off[s[k] - 1]; const int64_t* b = ids + off[s[k]];
                const int64_t* it = std::lower_bound(a, b, (int64_t)s[q]);
                if (it == b || *it != s[q]) continue;
                double m[NM];
                vertex_flag_moments<D>(xs, n, &ps, s, k, m, q);
                double* o = out + (it - ids) * (1 + D);
                const double* x = xs + (s[k] - 1) * D;
                o[0] += m[0] / fact;
                for (int c = 0; c < D; ++c) o[1 + c] += (x[c] * m[0] + m[1 + c]) / fact;
            }
        }
    }
}

// convex hull by the facet walk of hvb_hull.cuh, driven sequentially (unbounded domain)
struct HullResult { int d; int64_t nf; std::vector<int64_t> facet; std::vector<double> normal, centre; int64_t raycasts, records, rounds, degenerate; };
template <int D>
static HullResult* run_hull(int64_t n, const double* xs, int ppc) {
    Dev<D> dv;
    memset(&dv, 0, sizeof(dv));
    dv.n = (int)n;
    double blo[D], bhi[D];
    for (int k = 0; k < D; ++k) { blo[k] = 1e300; bhi[k] = -1e300; }
    for (int64_t i = 0; i < n; ++i)
        for (int k = 0; k < D; ++k) { blo[k] = std::min(blo[k], xs[i * D + k]); bhi[k] = std::max(bhi[k], xs[i * D + k]); }
    int64_t ncell = setup_grid<D>(dv, blo, bhi, n, ppc > 0 ? ppc : default_points_per_cell(D));
    std::vector<int> cell(n), cstart(ncell + 1, 0), perm(n);
    for (int64_t i = 0; i < n; ++i) { cell[i] = cell_index<D>(dv, xs + i * D); cstart[cell[i] + 1]++; }
    for (int64_t c = 0; c < ncell; ++c) cstart[c + 1] += cstart[c];
    { std::vector<int> cur(cstart.begin(), cstart.end() - 1); for (int64_t i = 0; i < n; ++i) perm[cur[cell[i]]++] = (int)i; }
    std::vector<double> x64(n * D); std::vector<float> x32(n * X32<D>::STRIDE, 0.f);
    for (int64_t i = 0; i < n; ++i)
        for (int k = 0; k < D; ++k) { x64[i * D + k] = xs[(int64_t)perm[i] * D + k]; x32[i * X32<D>::STRIDE + k] = (float)(x64[i * D + k] - dv.lo[k]); }
    PlaneSet ps; memset(&ps, 0, sizeof(ps));
    std::vector<unsigned char> active(n, 1), hasv(n, 0);
    dv.cell_start = cstart.data(); dv.x32 = x32.data(); dv.x64 = x64.data(); dv.xcan = x64.data(); dv.planes = &ps; dv.active = active.data();
    dv.plane_tol = 1e-12; dv.t_min = 1e-12; dv.probe_scale = default_probe_scale(D); dv.probe_growth = 1e9; dv.fp32_filter = 1;
    int64_t vcap = estimate_vertices(D, n, 0);
    std::vector<int> vsig(vcap * (D + 1)); std::vector<double> vr(vcap * D);
    u32 vcount = 0;
    Counters ctr; memset(&ctr, 0, sizeof(ctr));
    dv.vsig = vsig.data(); dv.vr = vr.data(); dv.vcount = &vcount; dv.vcap = (u32)vcap; dv.has_vertex = hasv.data(); dv.ctr = &ctr;
    HullDev<D> hd;
    u32 fcap = (u32)vcap, fcount = 0;
    std::vector<int> fsig((size_t)fcap * D); std::vector<u32> fitem(fcap); std::vector<double> fu((size_t)fcap * D);
    u64 fts = next_pow2(2 * (u64)fcap), rts = next_pow2(2 * (u64)fcap * D);
    std::vector<u64> ftab(fts, 0), rtab(rts, 0);
    hd.fsig = fsig.data(); hd.fitem = fitem.data(); hd.fu = fu.data(); hd.fcount = &fcount; hd.fcap = fcap;
    hd.ftab = ftab.data(); hd.fmask = fts - 1; hd.rtab = rtab.data(); hd.rmask = rts - 1;
    u32 qcap = (u32)(2 * (u64)fcap * D);
    std::vector<u64> qa(qcap), qb(qcap);
    u32 na = 0, nb = 0;
    TileHost tile; LocalStats ls; memset(&ls, 0, sizeof(ls));
    // the extreme generator along axis 0 (search_max, chull.jl:244)
    int start = 0;
    for (int64_t i = 1; i < n; ++i) if (x64[i * D] > x64[(size_t)start * D]) start = (int)i;
    HullResult* R = new HullResult(); R->d = D; R->rounds = 0;
    bool ok = hull_seed<D, TileHost>(dv, hd, tile, start, 0, qa.data(), &na, qcap, ls);
    while (ok && na > 0) {
        nb = 0;
        for (u32 i = 0; i < na; ++i) hull_step<D, TileHost>(dv, hd, tile, qa[i], qb.data(), &nb, qcap, ls);
        qa.swap(qb); na = nb; ++R->rounds;
    }
    R->raycasts = ls.raycasts; R->records = vcount; R->degenerate = ls.degenerate + ((ctr.flags & FLAG_OVERFLOW_MASK) ? 1000000 : 0);
    R->nf = 0;
    for (u32 f = 0; f < fcount; ++f) {
        if (fsig[(size_t)f * D] < 0) continue;
        std::vector<int64_t> e;
        for (int k = 0; k < D; ++k) e.push_back((int64_t)perm[fsig[(size_t)f * D + k]] + 1);
        std::sort(e.begin(), e.end());
        R->facet.insert(R->facet.end(), e.begin(), e.end());
        for (int k = 0; k < D; ++k) R->normal.push_back(fu[(size_t)f * D + k]);
        ++R->nf;
    }
    return R;
}

// convex hull by gift wrapping (hvb_wrap.cuh), driven sequentially: per round prepare -> scan -> commit, as the three kernels
// of the device do; `slots` cuts the generators into that many pieces whose partial results are merged (the device's chunks)
template <int D>
static HullResult* run_wrap(int64_t n, const double* xs, int slots, int fp32) {
    Dev<D> dv;
    memset(&dv, 0, sizeof(dv));
    dv.n = (int)n;
    double blo[D], bhi[D];
    for (int k = 0; k < D; ++k) { blo[k] = 1e300; bhi[k] = -1e300; }
    for (int64_t i = 0; i < n; ++i)
        for (int k = 0; k < D; ++k) { blo[k] = std::min(blo[k], xs[i * D + k]); bhi[k] = std::max(bhi[k], xs[i * D + k]); }
    int64_t ncell = setup_grid<D>(dv, blo, bhi, n, default_points_per_cell(D));
    std::vector<int> cell(n), cstart(ncell + 1, 0), perm(n);
    for (int64_t i = 0; i < n; ++i) { cell[i] = cell_index<D>(dv, xs + i * D); cstart[cell[i] + 1]++; }
    for (int64_t c = 0; c < ncell; ++c) cstart[c + 1] += cstart[c];
    { std::vector<int> cur(cstart.begin(), cstart.end() - 1); for (int64_t i = 0; i < n; ++i) perm[cur[cell[i]]++] = (int)i; }
    std::vector<double> x64(n * D); std::vector<float> x32((n + 8) * X32<D>::STRIDE, 0.f);
    for (int64_t i = 0; i < n; ++i)
        for (int k = 0; k < D; ++k) { x64[i * D + k] = xs[(int64_t)perm[i] * D + k]; x32[i * X32<D>::STRIDE + k] = (float)(x64[i * D + k] - dv.lo[k]); }
    Counters ctr; memset(&ctr, 0, sizeof(ctr));
    dv.x32 = x32.data(); dv.x64 = x64.data(); dv.xcan = x64.data(); dv.ctr = &ctr;
    HullDev<D> hd;
    u32 fcap = (u32)std::max<int64_t>(1 << 12, 64 * n), fcount = 0;
    std::vector<int> fsig((size_t)fcap * D); std::vector<u32> fitem(fcap); std::vector<double> fu((size_t)fcap * D);
    u64 fts = next_pow2(2 * (u64)fcap), rts = next_pow2(2 * (u64)fcap * D);
    std::vector<u64> ftab(fts, 0), rtab(rts, 0);
    hd.fsig = fsig.data(); hd.fitem = fitem.data(); hd.fu = fu.data(); hd.fcount = &fcount; hd.fcap = fcap;
    hd.ftab = ftab.data(); hd.fmask = fts - 1; hd.rtab = rtab.data(); hd.rmask = rts - 1;
    WrapDev<D> wd;
    memset(&wd, 0, sizeof(wd));
    u32 qcap = fcap * D;
    std::vector<u64> qa(qcap), qb(qcap);
    std::vector<WrapQuery<D> > wq(qcap);
    u32 qcount[2] = {0, 0}, nqv[2] = {0, 0};
    WrapSeed seed[16]; memset(seed, 0, sizeof(seed));
    wd.wq = wq.data(); wd.wq_cap = qcap; wd.q[0] = qa.data(); wd.q[1] = qb.data(); wd.qcap = qcap; wd.qcount = qcount; wd.nq = nqv; wd.seed = seed;
    wrap_tolerances<D>(dv.ext, wd.E32, wd.tinyA);
    if (!fp32) wd.E32 = INFINITY;                 // every generator goes through the FP64 evaluation
    LocalStats ls; memset(&ls, 0, sizeof(ls));
    // the extreme generators along the axes (search_max, chull.jl:244): 2 D first facets
    for (int a = 0; a < 2 * D; ++a) {
        const int axis = a >> 1; const double sg = (a & 1) ? -1.0 : 1.0;
        int start = 0;
        for (int64_t i = 1; i < n; ++i) if (sg * x64[i * D + axis] > sg * x64[(size_t)start * D + axis]) start = (int)i;
        seed[a].ids[0] = start; seed[a].cnt = 1;
        for (int k = 0; k < D; ++k) seed[a].u[k] = (k == axis) ? sg : 0.0;
        qa[a] = WRAP_SEED_ENTRY + (u64)a;
    }
    qcount[0] = 2 * D;
    HullResult* R = new HullResult(); R->d = D; R->rounds = 0; R->raycasts = 0;
    int cur = 0;
    if (slots < 1) slots = 1;
    while (qcount[cur] > 0 && R->rounds < 1000000) {
        const int nxt = 1 - cur;
        qcount[nxt] = 0; nqv[cur] = 0;
        for (u32 i = 0; i < qcount[cur]; ++i) {
            WrapQuery<D> w;
            if (wrap_prepare<D>(dv, hd, wd, wd.q[cur][i], w, ls)) wq[nqv[cur]++] = w;
        }
        for (u32 q = 0; q < nqv[cur]; ++q) {
            double c1 = -INFINITY, c2 = -INFINITY; int g = -1;
            const int64_t chunk = (((n + slots - 1) / slots) + 3) & ~3LL;
            for (int s = 0; s < slots; ++s) {
                const int64_t p0 = s * chunk, p1 = std::min<int64_t>(p0 + chunk, n);
                WrapLane<D> lane;
                wrap_lane_init<D>(wq[q], lane);
                if (p1 > p0) wrap_scan<D>(dv, wq[q], x32.data() + (size_t)p0 * X32<D>::STRIDE, (int)p0, (int)(p1 - p0), wd.tinyA, wd.E32, lane, ls);
                wrap_settle<D>(dv, wq[q], wd.tinyA, wd.E32, lane, ls);
                wrap_merge(c1, g, c2, lane.c1, lane.id1, lane.c2);
            }
            wrap_commit<D>(dv, hd, wd, wq[q], c1, g, c2, nxt, ls);
            ++R->raycasts;
        }
        cur = nxt; ++R->rounds;
    }
    R->records = ls.cand64; R->degenerate = ls.degenerate + ls.seed_fail + ((ctr.flags & FLAG_OVERFLOW_MASK) ? 1000000 : 0);
    R->nf = 0;
    for (u32 f = 0; f < fcount; ++f) {
        if (fsig[(size_t)f * D] < 0) continue;
        std::vector<std::pair<int64_t, int> > e;
        for (int k = 0; k < D; ++k) e.push_back(std::make_pair((int64_t)perm[fsig[(size_t)f * D + k]] + 1, fsig[(size_t)f * D + k]));
        std::sort(e.begin(), e.end());
        double P[D][D], nrm[D], cen[D];
        for (int i = 0; i < D; ++i) for (int k = 0; k < D; ++k) P[i][k] = x64[(size_t)e[i].second * D + k];
        if (!wrap_facet_geometry<D>(P, &fu[(size_t)f * D], nrm, cen)) R->degenerate += 1;
        for (int k = 0; k < D; ++k) R->facet.push_back(e[k].first);
        for (int k = 0; k < D; ++k) R->normal.push_back(nrm[k]);
        for (int k = 0; k < D; ++k) R->centre.push_back(cen[k]);
        ++R->nf;
    }
    return R;
}

extern "C" {
void* hostsim_wrap(int dim, int64_t n, const double* xs, int slots, int fp32) {
    switch (dim) {
        case 2: return run_wrap<2>(n, xs, slots, fp32);
        case 3: return run_wrap<3>(n, xs, slots, fp32);
        case 4: return run_wrap<4>(n, xs, slots, fp32);
        case 5: return run_wrap<5>(n, xs, slots, fp32);
        case 6: return run_wrap<6>(n, xs, slots, fp32);
    }
    return 0;
}
void hostsim_wrap_centres(void* h, double* centre) { HullResult* R = (HullResult*)h; memcpy(centre, R->centre.data(), R->centre.size() * 8); }
void* hostsim_hull(int dim, int64_t n, const double* xs, int ppc) {
    switch (dim) {
        case 2: return run_hull<2>(n, xs, ppc);
        case 3: return run_hull<3>(n, xs, ppc);
        case 4: return run_hull<4>(n, xs, ppc);
        case 5: return run_hull<5>(n, xs, ppc);
        case 6: return run_hull<6>(n, xs, ppc);
    }
    return 0;
}
void hostsim_hull_counts(void* h, int64_t* out /*5*/) { HullResult* R = (HullResult*)h; out[0] = R->nf; out[1] = R->raycasts; out[2] = R->records; out[3] = R->rounds; out[4] = R->degenerate; }
void hostsim_hull_fetch(void* h, int64_t* facet, double* normal) { HullResult* R = (HullResult*)h; memcpy(facet, R->facet.data(), R->facet.size() * 8); memcpy(normal, R->normal.data(), R->normal.size() * 8); }
void hostsim_hull_free(void* h) { delete (HullResult*)h; }
void* hostsim_run(int dim, int64_t n, const double* xs, int P, const double* base, const double* normal,
                  int ppc, double probe_scale, int fp32, int seed_stride) {
    switch (dim) {
        case 2: return run<2>(n, xs, P, base, normal, ppc, probe_scale, fp32, seed_stride);
        case 3: return run<3>(n, xs, P, base, normal, ppc, probe_scale, fp32, seed_stride);
        case 4: return run<4>(n, xs, P, base, normal, ppc, probe_scale, fp32, seed_stride);
        case 5: return run<5>(n, xs, P, base, normal, ppc, probe_scale, fp32, seed_stride);
        case 6: return run<6>(n, xs, P, base, normal, ppc, probe_scale, fp32, seed_stride);
    }
    return 0;
}
void hostsim_counts(void* h, int64_t* nv, int64_t* nr, int64_t* ctr /*12*/) {
    SimResult* R = (SimResult*)h; *nv = R->nv; *nr = R->nr;
    const Counters& c = R->ctr;
    int64_t v[12] = {(int64_t)c.raycasts, (int64_t)c.dup_hits, (int64_t)c.closed_skips, (int64_t)c.cand32, (int64_t)c.cand64, (int64_t)c.rows,
                     (int64_t)c.stages, (int64_t)c.seeds, (int64_t)c.degenerate, (int64_t)c.seed_fail, (int64_t)c.dead, (int64_t)R->rounds};
    memcpy(ctr, v, sizeof(v));
}
void hostsim_fetch(void* h, int64_t* sig, double* r, int64_t* ray_edge) {
    SimResult* R = (SimResult*)h;
    if (!R->sig.empty()) { memcpy(sig, R->sig.data(), R->sig.size() * 8); memcpy(r, R->r.data(), R->r.size() * 8); }
    if (!R->ray_edge.empty()) memcpy(ray_edge, R->ray_edge.data(), R->ray_edge.size() * 8);
}
void hostsim_free(void* h) { delete (SimResult*)h; }
void hostsim_set_trace(int on) { g_want_trace = on != 0; }
void hostsim_set_active(int64_t n, const unsigned char* active) { g_active.assign(active, active + (active ? n : 0)); }
int64_t hostsim_trace_size(void* h) { return (int64_t)((SimResult*)h)->trace.size(); }
void hostsim_trace_fetch(void* h, int* out) { SimResult* R = (SimResult*)h; memcpy(out, R->trace.data(), R->trace.size() * sizeof(int)); }

void hostsim_areas(int dim, int64_t n, const double* xs, int P, const double* base, const double* normal, int64_t nv, const int64_t* sig,
                   const int64_t* off, const int64_t* ids, double* area) {
    switch (dim) {
        case 2: areas<2>(n, xs, P, base, normal, nv, sig, off, ids, area); break;
        case 3: areas<3>(n, xs, P, base, normal, nv, sig, off, ids, area); break;
        case 4: areas<4>(n, xs, P, base, normal, nv, sig, off, ids, area); break;
        case 5: areas<5>(n, xs, P, base, normal, nv, sig, off, ids, area); break;
        case 6: areas<6>(n, xs, P, base, normal, nv, sig, off, ids, area); break;
    }
}
void hostsim_area_moments(int dim, int64_t n, const double* xs, int P, const double* base, const double* normal, int64_t nv, const int64_t* sig,
                          const int64_t* off, const int64_t* ids, double* out) {
    switch (dim) {
        case 2: area_moments<2>(n, xs, P, base, normal, nv, sig, off, ids, out); break;
        case 3: area_moments<3>(n, xs, P, base, normal, nv, sig, off, ids, out); break;
        case 4: area_moments<4>(n, xs, P, base, normal, nv, sig, off, ids, out); break;
        case 5: area_moments<5>(n, xs, P, base, normal, nv, sig, off, ids, out); break;
        case 6: area_moments<6>(n, xs, P, base, normal, nv, sig, off, ids, out); break;
    }
}
void* hostsim_resolve(int dim, int64_t n, const double* xs, int P, const double* base, const double* normal) {
    switch (dim) {
        case 2: return run_resolve<2>(n, xs, P, base, normal);
        case 3: return run_resolve<3>(n, xs, P, base, normal);
        case 4: return run_resolve<4>(n, xs, P, base, normal);
        case 5: return run_resolve<5>(n, xs, P, base, normal);
        case 6: return run_resolve<6>(n, xs, P, base, normal);
    }
    return 0;
}
void hostsim_resolve_counts(void* h, int64_t* out /*5*/) {
    ResolveResult* R = (ResolveResult*)h;
    out[0] = (int64_t)R->off.size() - 1; out[1] = (int64_t)R->ids.size(); out[2] = (int64_t)R->nb_ids.size(); out[3] = R->maxlen; out[4] = R->nsimplicial;
}
void hostsim_resolve_fetch(void* h, int64_t* off, int64_t* ids, double* r, int64_t* nb_off, int64_t* nb_ids) {
    ResolveResult* R = (ResolveResult*)h;
    memcpy(off, R->off.data(), R->off.size() * 8); memcpy(ids, R->ids.data(), R->ids.size() * 8); memcpy(r, R->r.data(), R->r.size() * 8);
    memcpy(nb_off, R->nb_off.data(), R->nb_off.size() * 8); if (!R->nb_ids.empty()) memcpy(nb_ids, R->nb_ids.data(), R->nb_ids.size() * 8);
}
void hostsim_resolve_free(void* h) { delete (ResolveResult*)h; }
void hostsim_moments(int dim, int64_t n, const double* xs, int P, const double* base, const double* normal, int64_t nv, const int64_t* sig, double* out) {
    switch (dim) {
        case 2: moments<2>(n, xs, P, base, normal, nv, sig, out); break;
        case 3: moments<3>(n, xs, P, base, normal, nv, sig, out); break;
        case 4: moments<4>(n, xs, P, base, normal, nv, sig, out); break;
        case 5: moments<5>(n, xs, P, base, normal, nv, sig, out); break;
        case 6: moments<6>(n, xs, P, base, normal, nv, sig, out); break;
    }
}
void hostsim_volumes(int dim, int64_t n, const double* xs, int P, const double* base, const double* normal, int64_t nv, const int64_t* sig, double* vol) {
    switch (dim) {
        case 2: volumes<2>(n, xs, P, base, normal, nv, sig, vol); break;
        case 3: volumes<3>(n, xs, P, base, normal, nv, sig, vol); break;
        case 4: volumes<4>(n, xs, P, base, normal, nv, sig, vol); break;
        case 5: volumes<5>(n, xs, P, base, normal, nv, sig, vol); break;
        case 6: volumes<6>(n, xs, P, base, normal, nv, sig, vol); break;
    }
}
}
