// hvb_core.cuh -- the raycast vertex search, written once for the device (sm_100a) and, for debugging in a
// container without a GPU, compilable as plain host C++ (tests/hostsim).  Nothing here is a port of the Julia
// code: the reference's cell-ordered, pointer-chasing walk (sysvoronoi.jl:384-525) is re-designed as an
// edge frontier whose entries are solved by a tile of G cooperating lanes.
//
//   reference                                        here
//   ---------------------------------------------    ---------------------------------------------------------
//   u_qr (tools.jl:773-790)                          ortho_direction(): MGS2 in registers, sign by construction
//   raycast_des2(::HPUnion) (raycast.jl:794-970)     min_t_query(): staged probe balls over a uniform grid,
//     = 2.6 KD nn + 1 inrange                          FP32 filter with an explicit error bound + FP64 verify
//   mirrors via ExtendedNodes (extended.jl,          analytic plane candidates t = (off - n.r)/(n.u), id N+p
//     raycast.jl:354-375)
//   EdgeHashTable per cell (edgehashing.jl:66-111)   one global edge table (64-bit slots, CAS), closed bit
//   HeapDataBase + QueueHashTable (hvdatabase.jl)    vertex set: 64-bit slots {fingerprint, index} + SoA records
//   descent (raycast.jl:45-109)                      seed_item(): d successive min-t queries
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define HVB_HD __host__ __device__ __forceinline__
#define HVB_D __device__ __forceinline__
#else
#define HVB_HD inline
#define HVB_D inline
#endif

// trace hook of the host build (tests/hostsim records per-ray row / point counts for tools/simt_model.py); nothing on the device
#ifndef HVB_TRACE_EVENT
#define HVB_TRACE_EVENT(kind, val)
#endif

#ifndef HVB_MAX_PLANES
#define HVB_MAX_PLANES 32
#endif

namespace hvb {

typedef unsigned long long u64;
typedef unsigned int u32;

// ------------------------------------------------------------------------------------------------------------
// memory helpers: every structure that is written during a launch is read with ld.global.cg (L1 is not coherent)
// ------------------------------------------------------------------------------------------------------------
HVB_HD u64 ld_cg(const u64* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
HVB_HD int ld_cg(const int* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
HVB_HD double ld_cg(const double* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
HVB_HD u64 atom_cas(u64* p, u64 cmp, u64 val) {
#if defined(__CUDA_ARCH__)
    return atomicCAS(p, cmp, val);
#else
    u64 old = *p; if (old == cmp) *p = val; return old;
#endif
}
HVB_HD u32 atom_add(u32* p, u32 v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    u32 old = *p; *p += v; return old;
#endif
}
HVB_HD void atom_or(u64* p, u64 v) {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
HVB_HD void atom_or(u32* p, u32 v) {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
HVB_HD void mem_fence() {
#if defined(__CUDA_ARCH__)
    __threadfence();
#endif
}

// ------------------------------------------------------------------------------------------------------------
// tiles: G lanes that solve one frontier entry together.  TileHost (G = 1) is the debugging stand-in.
// ------------------------------------------------------------------------------------------------------------
struct TileHost {
    static const int SIZE = 1;
    HVB_HD int lane() const { return 0; }
    HVB_HD double shfl_xor(double v, int) const { return v; }
    HVB_HD int shfl_xor(int v, int) const { return v; }
    HVB_HD float shfl_xor(float v, int) const { return v; }
    HVB_HD int shfl(int v, int) const { return v; }
    HVB_HD u32 shfl(u32 v, int) const { return v; }
    HVB_HD u64 shfl(u64 v, int) const { return v; }
    HVB_HD u32 ballot(bool p) const { return p ? 1u : 0u; }
    HVB_HD void sync() const {}
};
#if defined(__CUDACC__)
template <int G>
struct TileDev {
    static const int SIZE = G;
    unsigned mask;
    int ln;
    __device__ __forceinline__ TileDev() {
        int wl = threadIdx.x & 31;
        ln = wl & (G - 1);
        mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl & ~(G - 1)));
    }
    __device__ __forceinline__ int lane() const { return ln; }
    __device__ __forceinline__ double shfl_xor(double v, int m) const { return G == 1 ? v : __shfl_xor_sync(mask, v, m, G); }
    __device__ __forceinline__ int shfl_xor(int v, int m) const { return G == 1 ? v : __shfl_xor_sync(mask, v, m, G); }
    __device__ __forceinline__ float shfl_xor(float v, int m) const { return G == 1 ? v : __shfl_xor_sync(mask, v, m, G); }
    __device__ __forceinline__ int shfl(int v, int src) const { return G == 1 ? v : __shfl_sync(mask, v, src, G); }
    __device__ __forceinline__ u32 shfl(u32 v, int src) const { return G == 1 ? v : __shfl_sync(mask, v, src, G); }
    __device__ __forceinline__ u64 shfl(u64 v, int src) const { return G == 1 ? v : __shfl_sync(mask, v, src, G); }
    // reconverges the lanes of this tile (they may still be split after lane-dependent work)
    __device__ __forceinline__ void sync() const { if (G > 1) __syncwarp(mask); }
    // ballot restricted to this tile, bit i = lane i of the tile
    __device__ __forceinline__ u32 ballot(bool p) const {
        u32 b = __ballot_sync(mask, p);
        return (G == 32) ? b : ((b & mask) >> ((threadIdx.x & 31) & ~(G - 1)));
    }
};
#endif

// ------------------------------------------------------------------------------------------------------------
// device-resident problem description
// ------------------------------------------------------------------------------------------------------------
struct PlaneSet {                      // boundary.jl:22-29 : outward unit normal n, offset = n.base ; inside: n.y <= off
    int P;
    int pad;
    double normal[HVB_MAX_PLANES * 6];
    double off[HVB_MAX_PLANES];
};

struct Counters {
    u64 raycasts, dup_hits, closed_skips, cand32, cand64, rows, stages, seeds, degenerate, seed_fail, dead;
    u32 flags;                         // bit0 vertex store full, bit1 frontier full, bit2 ray list full
    u32 pad;
};
enum { FLAG_VFULL = 1, FLAG_QFULL = 2, FLAG_RFULL = 4, FLAG_PAIRFULL = 8, FLAG_DEGEN = 16 };
const u32 FLAG_OVERFLOW_MASK = FLAG_VFULL | FLAG_QFULL | FLAG_RFULL;

template <int D>
struct Dev {
    // ---- read-only during the search -----------------------------------------------------------------
    int n;                             // generators
    double lo[D], h[D], inv_h[D];      // uniform grid: origin, cell size, 1/cell size
    int g[D];                          // cells per axis (linear index: axis D-1 fastest)
    double hmin, diag;                 // smallest cell edge, diagonal of the bounding box
    double ext;                        // largest bounding-box extent: scale of the FP32 coordinate error
    float h32[D], inv_h32[D], ext32;   // (float) of h, inv_h, ext: the FP32 row geometry must not convert them per row
    int pad32;
    const int* cell_start;             // [ncells + 1]
    const float* x32;                  // [n][X32<D>::STRIDE]  coordinates minus lo, FP32, padded (filter)
    const double* x64;                 // [n][D]  coordinates, FP64 (verification)
    const double* xcan;                // [n][D]  coordinates the returned vertex coordinates are solved from: x64, except when the
                                       //         search runs on perturbed generators (non-general position resolved, hvb_ctx.cuh)
    const PlaneSet* planes;
    const unsigned char* active;       // [n] 1 = this context walks the edges of that cell (slab / Iter)
    double plane_tol;                  // raycast-types.jl:229
    double t_min;                      // smallest accepted ray parameter: plane_tol (raycast.jl:887-889), except on a perturbed cloud
                                       // (non-general position resolved): there the true winner of a walk inside a cospherical set
                                       // sits at t ~ 1e-9 of the extent and must not be refused
    double probe_scale;
    double probe_growth;               // radius factor from one probe stage to the next (2; the hull walk jumps to the half-space)
    int fp32_filter;
    // ---- written during the search -------------------------------------------------------------------
    int* vsig;                         // [vcap][D+1] sorted internal ids; vsig[v][0] = -1 marks a dead record
    double* vr;                        // [vcap][D]
    u32* vcount; u32 vcap;
    u64* vtab; u64 vmask;              // vertex set: slot = fingerprint << 32 | (index + 1)
    u64* etab; u64 emask;              // edge table: see edge_slot()
    unsigned char* has_vertex;         // [n]
    u32* ray_item; double* ray_u; u32* ray_count; u32 ray_cap;   // unbounded edges: v << 3 | k, direction
    Counters* ctr;
};

// linear grid cell of a point (axis D-1 fastest); the same expression is used when the index is built
template <int D>
HVB_HD int cell_index(const Dev<D>& dv, const double* x) {
    int idx = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double v = fmin(fmax((x[k] - dv.lo[k]) * dv.inv_h[k], 0.0), (double)(dv.g[k] - 1));
        idx = idx * dv.g[k] + (int)floor(v);
    }
    return idx;
}

// FP32 filter coordinates are padded to a vector-load friendly stride: 2, 4, 4, 8, 8 floats for d = 2..6
template <int D>
struct X32 { static const int STRIDE = (D == 2) ? 2 : (D <= 4) ? 4 : 8; };

template <int D>
HVB_HD void load_x32(const float* base, int idx, float (&x)[D]) {
#if defined(__CUDA_ARCH__)
    float t[8];
    if (D == 2) {
        float2 v = __ldg(reinterpret_cast<const float2*>(base) + idx);
        t[0] = v.x; t[1] = v.y;
    } else if (D <= 4) {
        float4 v = __ldg(reinterpret_cast<const float4*>(base) + idx);
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    } else {
        float4 v = __ldg(reinterpret_cast<const float4*>(base) + 2 * (size_t)idx);
        float4 w = __ldg(reinterpret_cast<const float4*>(base) + 2 * (size_t)idx + 1);
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w; t[4] = w.x; t[5] = w.y; t[6] = w.z; t[7] = w.w;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = t[k];
#else
    for (int k = 0; k < D; ++k) x[k] = base[(size_t)idx * X32<D>::STRIDE + k];
#endif
}

struct LocalStats {
    u32 raycasts, dup_hits, closed_skips, cand32, cand64, rows, stages, seeds, degenerate, seed_fail, dead;
};

// ------------------------------------------------------------------------------------------------------------
// hashing of sorted id tuples (replaces fnv1a_hash, tools.jl:11-28; full keys are always compared)
// ------------------------------------------------------------------------------------------------------------
HVB_HD u64 mix64(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
template <int D>
HVB_HD u64 hash_ids(const int* ids, int count, int skip) {
    u64 h = 0x9e3779b97f4a7c15ULL;
#pragma unroll
    for (int i = 0; i < D + 1; ++i)
        if (i < count && i != skip) h = (h ^ (u64)(u32)ids[i]) * 0x100000001b3ULL + 0x632be59bd9b4e019ULL;
    return mix64(h);
}

// edge slot: [63] valid, [62:36] fingerprint, [35] closed, [34:3] vertex index, [2:0] dropped position
HVB_HD u64 edge_slot(u64 hash, u32 v, int k) {
    return (1ULL << 63) | (((hash >> 37) & 0x7ffffffULL) << 36) | ((u64)v << 3) | (u64)k;
}
const u64 EDGE_CLOSED = 1ULL << 35;
const u64 EDGE_FPMASK = (0x7ffffffULL << 36) | (1ULL << 63);

// ------------------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------------------
HVB_HD double inv_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}
template <int D>
HVB_HD double dotD(const double* a, const double* b) {
    double s = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) s += a[k] * b[k];
    return s;
}

// Unit vector orthogonal to the rows of V selected by `mask` (overwritten by an orthonormal basis), obtained by
// projecting `v` off them (modified Gram-Schmidt, two passes).  Replaces u_qr (tools.jl:773-790): with
// v = -(x_dropped - x0) the orientation "away from the dropped generator" holds by construction.  Rows are
// addressed statically (mask predicates) so that V stays in registers.
template <int D>
HVB_HD bool ortho_direction(double (&V)[D + 1][D], unsigned mask, double (&v)[D]) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < D + 1; ++i) {
        if ((mask >> i) & 1u) {
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
                for (int j = 0; j < D + 1; ++j)
                    if (j < i && ((mask >> j) & 1u)) {
                        double s = dotD<D>(V[i], V[j]);
#pragma unroll
                        for (int k = 0; k < D; ++k) V[i][k] -= s * V[j][k];
                    }
                double nr2 = dotD<D>(V[i], V[i]);
                ok &= (nr2 > 0);
                double inv = inv_sqrt(nr2);
#pragma unroll
                for (int k = 0; k < D; ++k) V[i][k] *= inv;
            }
        }
    }
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
        for (int j = 0; j < D + 1; ++j)
            if ((mask >> j) & 1u) {
                double s = dotD<D>(v, V[j]);
#pragma unroll
                for (int k = 0; k < D; ++k) v[k] -= s * V[j][k];
            }
        double nr2 = dotD<D>(v, v);
        ok &= (nr2 > 0);
        double inv = inv_sqrt(nr2);
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] *= inv;
    }
    return ok;
}

// ------------------------------------------------------------------------------------------------------------
// the min-t query
// ------------------------------------------------------------------------------------------------------------
template <int D>
struct RayQ {
    double r[D], u[D], x0[D];
    double c;          // half-space threshold on u.x : candidates need u.x > c   (raycast.jl:802-805, myskips :385)
    double R0sq;       // |x0 - r|^2
    double a;          // u.(x0 - r)
    double rnorm;      // |r| (scale of the reference's rounding estimate, get_t_hp_ raycast.jl:406)
    int excl[D + 1];   // ids that may not win (the origin vertex's generators)
    int nexcl;
};

struct Best {
    double t;          // smallest ray parameter so far
    int id;            // its generator (>= n : boundary plane id - n), -1 none
    double t2;         // runner-up, for the general-position check
    double tw;         // the reference's tie window on t for this winner (raycast.jl:902), see tie_window()
};

// The reference treats every candidate with t <= t_min + min(1e-7, (10 + d) * full_error) as tied with the winner
// (raycast.jl:902), full_error = (t * 1e-14 + |r| * 1e-15) / (u . normalize(x - x0)) (get_t_hp_, raycast.jl:399-407):
// the window widens for grazing candidates.  `den` = u . (x - x0), `dx2` = |x - x0|^2.
HVB_HD double tie_window(int D, double t, double rnorm, double den, double dx2) {
    const double w = (10.0 + D) * (t * 1e-14 + rnorm * 1e-15) * sqrt(dx2) / den;
    return fmin(1e-7, w);
}

// the window is only needed for a candidate that becomes the winner: it is computed inside that branch (an FP64 square
// root and a division per verified candidate otherwise)
HVB_HD void best_offer(Best& b, double t, int id, int D, double rnorm, double den, double dx2) {
    if (t < b.t || (t == b.t && id < b.id)) { b.t2 = b.t; b.t = t; b.id = id; b.tw = tie_window(D, t, rnorm, den, dx2); }
    else if (id != b.id && t < b.t2) b.t2 = t;
}
HVB_HD void best_merge(Best& b, double t, int id, double t2, double tw) {
    if (t < b.t || (t == b.t && id < b.id)) {
        double o = (id != b.id) ? b.t : b.t2;
        b.t2 = fmin(o, fmin(b.t2, t2)); b.t = t; b.id = id; b.tw = tw;
    } else {
        double o = (id != b.id) ? t : t2;
        b.t2 = fmin(b.t2, fmin(o, t2));
    }
}
template <class T>
HVB_HD void best_reduce(const T& tile, Best& b) {
#pragma unroll
    for (int m = T::SIZE / 2; m >= 1; m >>= 1) {
        double ot = tile.shfl_xor(b.t, m);
        int oid = tile.shfl_xor(b.id, m);
        double ot2 = tile.shfl_xor(b.t2, m);
        double otw = tile.shfl_xor(b.tw, m);
        best_merge(b, ot, oid, ot2, otw);
    }
}
// non-general position in the reference's sense: the runner-up lies inside the winner's tie window (the reference would
// append it to the signature, raycast.jl:926-949), or within 1e-12 relative of it
HVB_HD bool near_tie(const Best& b, double R0sq) {
    return b.t2 - b.t <= fmax(1e-12 * fmax(b.t, sqrt(R0sq)), b.tw);
}

// FP64 evaluation of one generator: get_t_hp (raycast.jl:427-432) under the predicate of myskips (:385)
template <int D>
HVB_HD bool verify64(const Dev<D>& dv, const RayQ<D>& q, int j, Best& best, LocalStats& ls) {
#pragma unroll
    for (int e = 0; e < D + 1; ++e)
        if (e < q.nexcl && q.excl[e] == j) return false;
    const double* x = dv.x64 + (size_t)j * D;
    double ux = 0, num = 0, den = 0, dx2 = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double xk = x[k];
        double dx = xk - q.x0[k];
        ux += q.u[k] * xk;
        num += dx * (q.x0[k] + xk - 2.0 * q.r[k]);
        den += q.u[k] * dx;
        dx2 += dx * dx;
    }
    ls.cand64++;
    if (!(ux > q.c) || !(den > 0)) return false;
    double t = num / (2.0 * den);
    if (!(t >= dv.t_min)) return false;               // raycast.jl:887-889
    best_offer(best, t, j, D, q.rnorm, den, dx2);
    return true;
}

// per-stage FP32 filter constants
struct Filt {
    float tb2;      // 2 * t_best, rounded up
    float en;       // bound on |num32 - num|
    float ed;       // bound on |den32 - den/2|
};
template <int D>
HVB_HD Filt make_filter(double t_best, double rho, double R0, double ext) {
    // coordinates are stored as fl32(x - lo) in [0, ext]: |q32 - q| <= eta per component
    const double eps = 5.9604644775390625e-08;        // 2^-24
    double eta = 3.0 * eps * ext;
    double Q = 2.0 * fmax(rho, R0), W = R0;
    double en = D * (2.0 * eta * (Q + W) + 2.0 * eps * Q * W + eta * eta) + (D + 3) * eps * D * Q * (Q + 2.0 * W);
    double ed = D * eta + (D + 2.0) * D * eps * Q;
    Filt f;
    double tb2 = 2.0 * t_best * (1.0 + 8.0 * eps);
    f.tb2 = (tb2 < 3.0e38) ? (float)tb2 * 1.0000005f : INFINITY;
    en *= 2.0; ed *= 2.0;
    f.en = (en < 3.0e38) ? (float)en * 1.0000005f : INFINITY;
    f.ed = (ed < 3.0e38) ? (float)ed * 1.0000005f : INFINITY;
    return f;
}

// Smallest t > 0 at which the ball centred at r + t u through x0 touches another generator or boundary plane.
// All lanes of the tile call this with identical q; all return the same Best.
// FP32 description of the current probe ball in grid-relative coordinates (used while the ball is of the order of
// the cloud; far-away / huge balls of unbounded edges keep the FP64 path below).  Every quantity carries the margin
// `m`, so the selected cells are a superset of the exact ones.
template <int D>
struct Ball32 {
    float cen[D];      // centre minus grid origin
    float rho2;        // squared radius, rounded up
    float m;           // absolute safety margin on coordinates
    float re[D];       // 1 / (cells of the box along axis k)
};

template <int D>
HVB_HD bool row_range32(const Dev<D>& dv, const float (&uf)[D], const float (&x0f)[D], const int (&clo)[D], const int (&chi)[D],
                        const Ball32<D>& b, int j, int& pa, int& pb) {
    int base = 0;
    float d2 = 0.f, umax = 0.f, uabs = 0.f;
    int rem = j;
    int cc[D];
#pragma unroll
    for (int k = D - 2; k >= 0; --k) {
        int e = chi[k] - clo[k] + 1;
        int qd = (int)(((float)rem + 0.5f) * b.re[k]);      // rem / e without an integer division (rem < 2^22)
        cc[k] = clo[k] + (rem - qd * e);
        rem = qd;
    }
#pragma unroll
    for (int k = 0; k < D - 1; ++k) {
        float hk = dv.h32[k];
        float blo = (float)cc[k] * hk - b.m;
        float bhi = blo + hk + 2.f * b.m;
        float dd = fmaxf(0.f, fmaxf(blo - b.cen[k], b.cen[k] - bhi));
        d2 = fmaf(dd, dd, d2);
        umax += fmaxf(uf[k] * (blo - x0f[k]), uf[k] * (bhi - x0f[k]));
        uabs += fabsf(uf[k]);
        base = base * dv.g[k] + cc[k];
    }
    if (!(d2 <= b.rho2)) return false;
    const int L = D - 1;
    float s = sqrtf(b.rho2 - d2) * 1.000001f + b.m;
    float zlo = b.cen[L] - s, zhi = b.cen[L] + s;
    float ul = uf[L];
    float slack = b.m * (uabs + fabsf(ul)) + 1e-5f * fabsf(umax) + 1e-6f * dv.ext32;
    if (ul > 1e-3f) zlo = fmaxf(zlo, x0f[L] - (umax + slack) / ul * 1.00001f - b.m);
    else if (ul < -1e-3f) zhi = fminf(zhi, x0f[L] - (umax + slack) / ul * 1.00001f + b.m);
    else if (umax + slack + fabsf(ul) * dv.ext32 * 2.f <= 0.f) return false;
    float ihl = dv.inv_h32[L];
    float gl = (float)dv.g[L];
    float vlo = fminf(fmaxf(zlo * ihl - 2e-3f, 0.f), gl - 1.f);
    float vhi = fminf(fmaxf(zhi * ihl + 2e-3f, -1.f), gl - 1.f);
    int z0 = (int)floorf(vlo), z1 = (int)floorf(vhi);
    if (z1 < z0) return false;
    const int* cs = dv.cell_start + (size_t)base * dv.g[L];
    pa = cs[z0]; pb = cs[z1 + 1];
    return true;
}

// Same as row_range32, but tells how far the row index may jump when the row is pruned: the rows of the cell box are
// numbered with axis D-2 fastest, so if the partial distance over the leading axes 0..k already exceeds the radius,
// every row that shares those leading cell coordinates is pruned at once.  Returns -1 if the row is usable (pa, pb
// set), else the index of the first row that is not pruned by the same test (> j).  Pays off for d >= 4, where most
// rows of the box miss the ball.
template <int D>
HVB_HD int row_try32(const Dev<D>& dv, const float (&uf)[D], const float (&x0f)[D], const int (&clo)[D], const int (&chi)[D],
                     const Ball32<D>& b, int j, int& pa, int& pb) {
    int rem = j;
    int dg[D], pv[D];
    int place = 1;
#pragma unroll
    for (int k = D - 2; k >= 0; --k) {
        int e = chi[k] - clo[k] + 1;
        int qd = (int)(((float)rem + 0.5f) * b.re[k]);
        dg[k] = rem - qd * e;
        rem = qd;
        pv[k] = place;
        place *= e;
    }
    int base = 0, prefix = 0;
    float d2 = 0.f, umax = 0.f, uabs = 0.f;
#pragma unroll
    for (int k = 0; k < D - 1; ++k) {
        int e = chi[k] - clo[k] + 1;
        int c = clo[k] + dg[k];
        prefix = prefix * e + dg[k];
        float hk = dv.h32[k];
        float blo = (float)c * hk - b.m;
        float bhi = blo + hk + 2.f * b.m;
        float dd = fmaxf(0.f, fmaxf(blo - b.cen[k], b.cen[k] - bhi));
        d2 = fmaf(dd, dd, d2);
        if (!(d2 <= b.rho2)) return (prefix + 1) * pv[k];
        umax += fmaxf(uf[k] * (blo - x0f[k]), uf[k] * (bhi - x0f[k]));
        uabs += fabsf(uf[k]);
        base = base * dv.g[k] + c;
    }
    const int L = D - 1;
    float s = sqrtf(b.rho2 - d2) * 1.000001f + b.m;
    float zlo = b.cen[L] - s, zhi = b.cen[L] + s;
    float ul = uf[L];
    float slack = b.m * (uabs + fabsf(ul)) + 1e-5f * fabsf(umax) + 1e-6f * dv.ext32;
    if (ul > 1e-3f) zlo = fmaxf(zlo, x0f[L] - (umax + slack) / ul * 1.00001f - b.m);
    else if (ul < -1e-3f) zhi = fminf(zhi, x0f[L] - (umax + slack) / ul * 1.00001f + b.m);
    else if (umax + slack + fabsf(ul) * dv.ext32 * 2.f <= 0.f) return j + 1;
    float ihl = dv.inv_h32[L];
    float gl = (float)dv.g[L];
    float vlo = fminf(fmaxf(zlo * ihl - 2e-3f, 0.f), gl - 1.f);
    float vhi = fminf(fmaxf(zhi * ihl + 2e-3f, -1.f), gl - 1.f);
    int z0 = (int)floorf(vlo), z1 = (int)floorf(vhi);
    if (z1 < z0) return j + 1;
    const int* cs = dv.cell_start + (size_t)base * dv.g[L];
    pa = cs[z0]; pb = cs[z1 + 1];
    return -1;
}

// Point range [pa, pb) of grid row j of the current cell box: the cells of that row that can hold a generator
// inside the search ball and on the positive side of the edge's hyperplane.  Issues the two cell_start loads.
template <int D>
HVB_HD bool row_range(const Dev<D>& dv, const RayQ<D>& q, const int (&clo)[D], const int (&chi)[D], const double (&cen)[D],
                      double rho2, int j, int& pa, int& pb) {
    int base = 0;
    double d2 = 0, umax = 0;
    int rem = j;
    int cc[D];
#pragma unroll
    for (int k = D - 2; k >= 0; --k) {
        int e = chi[k] - clo[k] + 1;
        int qd = rem / e;
        cc[k] = clo[k] + (rem - qd * e);
        rem = qd;
    }
#pragma unroll
    for (int k = 0; k < D - 1; ++k) {
        double blo = dv.lo[k] + cc[k] * dv.h[k] - 1e-9 * dv.h[k];
        double bhi = blo + dv.h[k] * (1.0 + 2e-9);
        double dd = fmax(0.0, fmax(blo - cen[k], cen[k] - bhi));
        d2 += dd * dd;
        umax += fmax(q.u[k] * (blo - q.x0[k]), q.u[k] * (bhi - q.x0[k]));
        base = base * dv.g[k] + cc[k];
    }
    if (!(d2 <= rho2)) return false;
    const int L = D - 1;
    double s = sqrt(rho2 - d2);
    double zlo = cen[L] - s, zhi = cen[L] + s;
    double ul = q.u[L];
    double slack = 1e-9 * (dv.h[L] + fabs(umax));
    if (ul > 1e-300) zlo = fmax(zlo, q.x0[L] - (umax + slack) / ul);
    else if (ul < -1e-300) zhi = fmin(zhi, q.x0[L] - (umax + slack) / ul);
    else if (umax + slack <= 0) return false;
    double gl = (double)dv.g[L];
    double vlo = fmin(fmax((zlo - dv.lo[L]) * dv.inv_h[L] - 1e-9, 0.0), gl - 1.0);
    double vhi = fmin(fmax((zhi - dv.lo[L]) * dv.inv_h[L] + 1e-9, -1.0), gl - 1.0);
    int z0 = (int)floor(vlo), z1 = (int)floor(vhi);
    if (z1 < z0) return false;
    const int* cs = dv.cell_start + (size_t)base * dv.g[L];
    pa = cs[z0]; pb = cs[z1 + 1];
    return true;
}

// row_range with block pruning for the FP64 path (the huge balls and the half-spaces of unbounded edges, where the cell
// box is the whole grid): returns -1 if row j is usable (pa, pb set), else the first row index that is not pruned by the
// same test.  Two tests cut whole blocks of rows that share leading cell coordinates: the partial distance to the ball's
// centre (as row_try32), and the half-space: if the best u.(x - x0) over the leading cells so far plus the best the
// REMAINING axes can contribute anywhere in the grid (msuf) is not positive, no generator behind that cell prefix lies
// beyond the edge's hyperplane.  An unbounded edge -- a facet of the convex hull -- must certify exactly that: an empty
// half-space; without this test every such ray visited every row of the grid.
template <int D>
HVB_HD int row_try64(const Dev<D>& dv, const RayQ<D>& q, const int (&clo)[D], const int (&chi)[D], const double (&cen)[D],
                     double rho2, const double (&msuf)[D], int j, int& pa, int& pb) {
    int rem = j;
    int dg[D], pv[D];
    int place = 1;
#pragma unroll
    for (int k = D - 2; k >= 0; --k) {
        const int e = chi[k] - clo[k] + 1;
        const int qd = rem / e;
        dg[k] = rem - qd * e;
        rem = qd;
        pv[k] = place;
        place *= e;
    }
    const double slk = 4e-9 * dv.diag;
    int base = 0, prefix = 0;
    double d2 = 0, umax = 0;
#pragma unroll
    for (int k = 0; k < D - 1; ++k) {
        const int e = chi[k] - clo[k] + 1;
        const int c = clo[k] + dg[k];
        prefix = prefix * e + dg[k];
        const double blo = dv.lo[k] + c * dv.h[k] - 1e-9 * dv.h[k];
        const double bhi = blo + dv.h[k] * (1.0 + 2e-9);
        const double dd = fmax(0.0, fmax(blo - cen[k], cen[k] - bhi));
        d2 += dd * dd;
        if (!(d2 <= rho2)) return (prefix + 1) * pv[k];
        umax += fmax(q.u[k] * (blo - q.x0[k]), q.u[k] * (bhi - q.x0[k]));
        if (umax + msuf[k] + slk <= 0) return (prefix + 1) * pv[k];
        base = base * dv.g[k] + c;
    }
    const int L = D - 1;
    const double s = sqrt(rho2 - d2);
    double zlo = cen[L] - s, zhi = cen[L] + s;
    const double ul = q.u[L];
    const double slack = 1e-9 * (dv.h[L] + fabs(umax));
    if (ul > 1e-300) zlo = fmax(zlo, q.x0[L] - (umax + slack) / ul);
    else if (ul < -1e-300) zhi = fmin(zhi, q.x0[L] - (umax + slack) / ul);
    else if (umax + slack <= 0) return j + 1;
    const double gl = (double)dv.g[L];
    const double vlo = fmin(fmax((zlo - dv.lo[L]) * dv.inv_h[L] - 1e-9, 0.0), gl - 1.0);
    const double vhi = fmin(fmax((zhi - dv.lo[L]) * dv.inv_h[L] + 1e-9, -1.0), gl - 1.0);
    const int z0 = (int)floor(vlo), z1 = (int)floor(vhi);
    if (z1 < z0) return j + 1;
    const int* cs = dv.cell_start + (size_t)base * dv.g[L];
    pa = cs[z0]; pb = cs[z1 + 1];
    return -1;
}

// FP32 bookkeeping of a probe stage: the candidate with the smallest FP32 UPPER bound of 2t (`cb`) and at most one
// rival whose FP32 interval overlaps it (`cr`).  Everything else is decided in FP32; the FP64 evaluation
// (get_t_hp, raycast.jl:427) runs once per stage for cb -- all lanes of a warp reach it together -- and only in
// rare cases for a rival.  id < 0: empty.
struct Cand32 { int id; float lo, hi; };
struct ScanState {
    Cand32 cb, cr;
    bool tighten;          // FP32 upper bounds may be trusted for this ray (see min_t_query)
    bool rejected;         // FP64 refused a candidate that had tightened the bound: the stage is redone without tightening
};

HVB_HD int lowest_bit(unsigned m) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)m) - 1;
#else
    int i = 0; while (!((m >> i) & 1u)) ++i; return i;
#endif
}
HVB_HD float fast_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
HVB_HD float upper_2t(double t) { return (float)(2.0 * t * (1.0 + 4.8e-7)) * 1.0000005f; }

// FP32 pass over the points [pa, pb) in chunks of four: all loads of a chunk are issued before any arithmetic.
template <int D>
HVB_HD void scan_points(const Dev<D>& dv, const RayQ<D>& q, const float (&uf)[D], const float (&w2f)[D], const float (&x0f)[D],
                        Filt& flt, int pa, int pb, ScanState& st, Best& best, LocalStats& ls, bool new_row = true) {
    ls.rows += new_row ? 1u : 0u;
    ls.cand32 += (u32)(pb - pa);
    if (!dv.fp32_filter) {
        for (int p = pa; p < pb; ++p) verify64<D>(dv, q, p, best, ls);
        return;
    }
#ifndef HVB_SCAN_U
#define HVB_SCAN_U 4
#endif
    const int U = HVB_SCAN_U;
    for (int p = pa; p < pb; p += U) {
        float x[U][D];
#pragma unroll
        for (int i = 0; i < U; ++i) load_x32<D>(dv.x32, (p + i < pb) ? (p + i) : (pb - 1), x[i]);
        // only the two sums are kept per point; the four interval ends are rebuilt for the (few) survivors
        float nms[U], dens[U];
        unsigned pmask = 0;
#pragma unroll
        for (int i = 0; i < U; ++i) {
            float den = 0.f, nm = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                float qk = x[i][k] - x0f[k];
                den = fmaf(uf[k], qk, den);
                nm = fmaf(qk, qk - w2f[k], nm);
            }
            nms[i] = nm; dens[i] = den;
            const bool pass = (p + i < pb) && (den + flt.ed > 0.f) && (nm - flt.en <= flt.tb2 * (den + flt.ed));
            pmask |= pass ? (1u << i) : 0u;
        }
        // survivors are handled one per loop trip, whatever their position in the chunk: the lanes of a warp that have
        // a survivor run this (branchy) bookkeeping together instead of position by position
        while (pmask) {
            const int i = lowest_bit(pmask);
            pmask &= pmask - 1u;
            HVB_TRACE_EVENT(4, 0);
            float nm_i = nms[0], den_i = dens[0];
#pragma unroll
            for (int b = 1; b < U; ++b) { nm_i = (i == b) ? nms[b] : nm_i; den_i = (i == b) ? dens[b] : den_i; }
            const float nlo_i = nm_i - flt.en, nhi_i = nm_i + flt.en, dh_i = den_i + flt.ed, dl_i = den_i - flt.ed;
            if (!(nlo_i <= flt.tb2 * dh_i)) continue;                    // re-checked: the bound may have tightened within the chunk
            const int id = p + i;
            bool excluded = false;
#pragma unroll
            for (int e = 0; e < D + 1; ++e) excluded |= (e < q.nexcl && q.excl[e] == id);
            if (excluded) continue;
            // an FP32 upper bound exists when the denominator is safely positive and t is safely > 0: such a
            // candidate is valid in FP64 as well (u.x > c, den > 0, t >= plane_tol)
            const bool bounded = st.tighten && dl_i > flt.ed && nlo_i > 0.f;
            int to_verify = -1;
            if (!bounded) to_verify = id;
            else {
                // approximate division (<= 2 ulp) inside margins of 6.7 / 16.8 ulp
                float lo = fast_div(nlo_i, dh_i); lo -= fabsf(lo) * 4e-7f;
                float hi = fast_div(nhi_i, dl_i) * 1.000001f;
                if (hi < st.cb.hi) {
                    // new FP32 best; the old one stays as the rival if its interval still overlaps
                    if (st.cb.id >= 0 && st.cb.lo <= hi) {
                        if (st.cr.id >= 0 && st.cr.lo <= hi) to_verify = st.cr.id;    // no room: settle the displaced rival now
                        st.cr = st.cb;
                    } else if (st.cr.id >= 0 && st.cr.lo > hi) st.cr.id = -1;
                    st.cb.id = id; st.cb.lo = lo; st.cb.hi = hi;
                    flt.tb2 = fminf(flt.tb2, hi);
                } else if (st.cr.id < 0) { st.cr.id = id; st.cr.lo = lo; st.cr.hi = hi; }
                else to_verify = id;
            }
            if (to_verify >= 0) {
                double before = best.t;
                verify64<D>(dv, q, to_verify, best, ls);
                if (best.t < before) flt.tb2 = fminf(flt.tb2, upper_2t(best.t));
            }
        }
    }
}

// end of a probe stage: the FP64 evaluation of the FP32 winner (and of a surviving rival)
template <int D>
HVB_HD void settle_stage(const Dev<D>& dv, const RayQ<D>& q, ScanState& st, float tb2, Best& best, LocalStats& ls) {
    if (st.cb.id >= 0 && st.cb.lo <= tb2) {
        if (!verify64<D>(dv, q, st.cb.id, best, ls)) st.rejected = true;
    }
    if (st.cr.id >= 0 && st.cr.lo <= tb2) verify64<D>(dv, q, st.cr.id, best, ls);
    st.cb.id = -1; st.cb.hi = INFINITY; st.cb.lo = 0.f;
    st.cr.id = -1;
}

// boundary planes as candidates of a ray: mirror images of x0 (raycast.jl:354-375, extended.jl:131-140), analytically
template <int D>
HVB_HD void plane_candidates(const Dev<D>& dv, const RayQ<D>& q, Best& best) {
    const PlaneSet* ps = dv.planes;
    const int P = ps->P;
    double ux0 = dotD<D>(q.u, q.x0);
    for (int p = 0; p < P; ++p) {
        bool ex = false;
#pragma unroll
        for (int e = 0; e < D + 1; ++e) ex |= (e < q.nexcl && q.excl[e] == dv.n + p);
        if (ex) continue;
        const double* nrm = ps->normal + p * 6;
        double nu = 0, nx0 = 0, nr = 0;
#pragma unroll
        for (int k = 0; k < D; ++k) { nu += nrm[k] * q.u[k]; nx0 += nrm[k] * q.x0[k]; nr += nrm[k] * q.r[k]; }
        double s = ps->off[p] - nx0;                       // distance of x0 to the plane (> 0 inside)
        if (!(ux0 + 2.0 * s * nu > q.c) || !(nu > 0)) continue;
        double t = (ps->off[p] - nr) / nu;
        if (!(t >= dv.t_min)) continue;
        // a plane is the mirror image of x0: x - x0 = 2 s n, u . (x - x0) = 2 s nu
        best_offer(best, t, dv.n + p, D, q.rnorm, 2.0 * s * nu, 4.0 * s * s);
    }
}

// Smallest t > 0 at which the ball centred at r + t u through x0 touches another generator or boundary plane.
// All lanes of the tile call this with identical q; all return the same Best.
// Squared distance of x0 to the ray's line, and the ball centred at r + T u through x0 (radius, squared radius with the
// row selection's safety factor).  Both come from the part of x0 - r perpendicular to u: the textbook forms R0^2 - a^2
// and R0^2 - 2 T a + T^2 cancel catastrophically when the origin vertex lies far outside the cloud (a flat simplex on
// the hull of an unbounded cloud has R0 ~ 1e7 cloud diameters, and the ball shrinks to the size of the cloud with the
// first candidate): the rounding of R0^2 then exceeded the whole radius and rows that held the winner were pruned.
// The absolute term covers the rounding of r + T u, of a and of the perpendicular part (a few ulp of |r| + R0 + |T|).
template <int D>
HVB_HD double ray_perp2(const RayQ<D>& q) {
    double s = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) { const double p = (q.x0[k] - q.r[k]) - q.a * q.u[k]; s += p * p; }
    return s;
}
template <int D>
HVB_HD void ray_ball(const RayQ<D>& q, double perp2, double R0, double T, double& rho, double& rho2) {
    const double dT = T - q.a;
    rho = sqrt(perp2 + dT * dT) + 1.5e-14 * (q.rnorm + R0 + fabs(T));
    rho2 = rho * rho * (1.0 + 1e-12) + 1e-300;
}

// Shrinks the search ball of a probe stage to the best bound so far (tb2 / 2 bounds the winner once a candidate tightened
// it).  With the FP32 row geometry this is FP32 arithmetic on rounded-up quantities (the ball only selects cells, a superset
// is all it has to be); the filter constants of the larger ball stay valid bounds.
template <int D>
HVB_HD void shrink_ball(const Dev<D>& dv, const RayQ<D>& q, bool use32, const Best& best, Filt& flt, const float (&uf)[D],
                        const float (&r32)[D], float a32, float perp2f, float& Ts32, Ball32<D>& b32, double& Ts, double (&cen)[D],
                        double& rho2, double& rho, double R0, double perp2) {
    if (use32) {
        const float tsh = fminf((float)best.t * 1.0000002f, 0.5f * flt.tb2 * 1.000001f);
        if (tsh < 0.92f * Ts32) {
            Ts32 = tsh;
#pragma unroll
            for (int k = 0; k < D; ++k) b32.cen[k] = fmaf(Ts32, uf[k], r32[k]);
            const float dT = fabsf(Ts32 - a32) + 4e-7f * (fabsf(a32) + Ts32);
            b32.rho2 = fmaf(dT, dT, perp2f) * 1.00001f;
        }
    } else {
        const double Tshr = fmin(best.t, 0.5 * (double)flt.tb2);
        if (Tshr < 0.92 * Ts) {
            Ts = Tshr;
#pragma unroll
            for (int k = 0; k < D; ++k) cen[k] = q.r[k] + Ts * q.u[k];
            ray_ball<D>(q, perp2, R0, Ts, rho, rho2);
            { float keep = flt.tb2; flt = make_filter<D>(Ts, rho, R0, dv.ext); flt.tb2 = fminf(flt.tb2, keep); }
        }
    }
}

template <int D, class T>
HVB_HD Best min_t_query(const Dev<D>& dv, const T& tile, const RayQ<D>& q, LocalStats& ls) {
    Best best;
    best.t = INFINITY; best.id = -1; best.t2 = INFINITY; best.tw = 0.0;
    const int lane = tile.lane();
    ls.raycasts += (lane == 0);
    HVB_TRACE_EVENT(0, 0);

    plane_candidates<D>(dv, q, best);

    // ---- FP32 copies of the ray ---------------------------------------------------------------------------
    float uf[D], w2f[D], x0f[D], r32[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        uf[k] = (float)q.u[k];
        w2f[k] = (float)(2.0 * (q.r[k] - q.x0[k]));
        x0f[k] = (float)(q.x0[k] - dv.lo[k]);
        r32[k] = (float)(q.r[k] - dv.lo[k]);
    }
    const double R0 = sqrt(q.R0sq);
    const double R0p = fmax(R0, 0.5 * dv.hmin);
    const double perp2 = ray_perp2<D>(q);                     // squared distance of x0 to the ray's line
    float a32 = (float)q.a;
    float perp2f = (float)(perp2 * (1.0 + 1e-6)) * 1.000001f;
#if defined(__CUDA_ARCH__)
    // Without this the compiler re-derives the FP32 copies from the doubles at their uses inside the scan loops
    // (FP64 subtract + FP64->FP32 conversion per use: 12 % of the warp instructions at d = 5,
    // profiles/r1_srcprofile_k_walk.md): values that come out of an asm statement cannot be rematerialised.
#pragma unroll
    for (int k = 0; k < D; ++k) asm volatile("" : "+f"(uf[k]), "+f"(w2f[k]), "+f"(x0f[k]), "+f"(r32[k]));
    asm volatile("" : "+f"(a32), "+f"(perp2f));
#endif
    double scale = dv.probe_scale;
    ScanState st;
    st.cb.id = -1; st.cb.lo = 0.f; st.cb.hi = INFINITY; st.cr.id = -1; st.cr.lo = 0.f; st.cr.hi = INFINITY;
    st.rejected = false;
    st.tighten = true;

    for (int stage = 0; stage < 96; ++stage) {
        ls.stages += (lane == 0);
        double rho_t = scale * R0p;
        // a probe ball several times the size of the cloud selects what the half-space does: go there at once (it is exact)
        double Tst = (rho_t > 4.0 * dv.diag) ? INFINITY : q.a + sqrt(fmax(rho_t * rho_t - perp2, 0.0));
        double Ts = fmin(Tst, best.t);
        bool halfspace_mode = !(Ts < INFINITY);
        // current search ball
        double cen[D], rho2, rho;
        if (halfspace_mode) {
#pragma unroll
            for (int k = 0; k < D; ++k) cen[k] = q.r[k];
            rho = 1e150; rho2 = 1e300;
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) cen[k] = q.r[k] + Ts * q.u[k];
            ray_ball<D>(q, perp2, R0, Ts, rho, rho2);
        }
        // cell box of the ball
        int clo[D], chi[D];
        int nrows = 1;
        bool empty = false;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            double gk = (double)dv.g[k];
            double vlo = fmin(fmax((cen[k] - rho - dv.lo[k]) * dv.inv_h[k] - 1e-9, 0.0), gk - 1.0);
            double vhi = fmin(fmax((cen[k] + rho - dv.lo[k]) * dv.inv_h[k] + 1e-9, -1.0), gk - 1.0);
            clo[k] = (int)floor(vlo); chi[k] = (int)floor(vhi);
            if (chi[k] < clo[k]) empty = true;
            if (k < D - 1) nrows *= (chi[k] - clo[k] + 1);
        }
        if (empty) nrows = 0;
        HVB_TRACE_EVENT(3, nrows);
        Filt flt = make_filter<D>(halfspace_mode ? INFINITY : Ts, rho, R0, dv.ext);
        // FP32 upper bounds certify validity only while the filter's denominator margin dominates the half-space slack
        if (!((float)(fabs(q.c) * 8e-12) < flt.ed)) st.tighten = false;
        const double Ts0 = Ts;
        // FP32 row geometry while the ball is comparable to the cloud; FP64 for the huge balls of unbounded edges
        double cmax = 0;                      // the centres of the shrinking ball lie between r and cen
#pragma unroll
        for (int k = 0; k < D; ++k) cmax = fmax(cmax, fmax(fabs(cen[k] - dv.lo[k]), fabs(q.r[k] - dv.lo[k])));
        float Ts32 = halfspace_mode ? INFINITY : (float)Ts * 1.0000002f;
        const bool use32 = !halfspace_mode && rho < 32.0 * dv.diag && cmax < 32.0 * dv.diag && nrows < (1 << 22);
        Ball32<D> b32;
        if (use32) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                b32.cen[k] = (float)(cen[k] - dv.lo[k]);
                b32.re[k] = 1.0f / (float)(chi[k] - clo[k] + 1);
            }
            b32.m = (float)(2e-6 * (dv.ext + cmax + rho));
            b32.rho2 = (float)(rho2 * (1.0 + 1e-5)) * 1.000001f;
        }

        if (D >= 4 || !use32) {
            // d >= 4, and the FP64 path of every dimension: walk the rows with block pruning (row_try32 / row_try64); a lane
            // keeps its residue class lane mod G
            double msuf[D];                     // FP64 path: best u.(x - x0) the axes behind k can contribute anywhere in the grid
            if (!use32) {
                double acc = 0;
#pragma unroll
                for (int k = D - 1; k >= 0; --k) {
                    msuf[k] = acc;
                    const double glo = dv.lo[k] - 1e-9 * dv.h[k], ghi = dv.lo[k] + dv.g[k] * dv.h[k] * (1.0 + 1e-9);
                    acc += fmax(q.u[k] * (glo - q.x0[k]), q.u[k] * (ghi - q.x0[k]));
                }
            }
            int j = lane;
            int pa = 0, pb = 0;
            for (;;) {
                bool have = false;
                while (j < nrows) {
                    int jn = use32 ? row_try32<D>(dv, uf, x0f, clo, chi, b32, j, pa, pb)
                                   : row_try64<D>(dv, q, clo, chi, cen, rho2, msuf, j, pa, pb);
                    HVB_TRACE_EVENT(1, 0);
                    if (jn < 0) { have = true; j += T::SIZE; break; }
                    j = (T::SIZE == 1) ? jn : jn + ((T::SIZE - ((jn - lane) % T::SIZE)) % T::SIZE);
                }
                if (T::SIZE > 1) { if (tile.ballot(have) == 0u) break; }
                else if (!have) break;
                HVB_TRACE_EVENT(2, (have && pb > pa) ? pb - pa : 0);
                if (have && pb > pa) scan_points<D>(dv, q, uf, w2f, x0f, flt, pa, pb, st, best, ls);
                if (T::SIZE > 1) {
                    best_reduce(tile, best);
#pragma unroll
                    for (int m = T::SIZE / 2; m >= 1; m >>= 1) flt.tb2 = fminf(flt.tb2, tile.shfl_xor(flt.tb2, m));
                }
                shrink_ball<D>(dv, q, use32, best, flt, uf, r32, a32, perp2f, Ts32, b32, Ts, cen, rho2, rho, R0, perp2);
            }
        } else {
        // rows lane, lane + G, ... ; the next row's range is requested before the current one is scanned
        const int niter = (nrows + T::SIZE - 1) / T::SIZE;
        int pa = 0, pb = 0, pa_n = 0, pb_n = 0;
        bool have = (lane < nrows) && (use32 ? row_range32<D>(dv, uf, x0f, clo, chi, b32, lane, pa, pb)
                                             : row_range<D>(dv, q, clo, chi, cen, rho2, lane, pa, pb));
        for (int it = 0; it < niter; ++it) {
            int jn = (it + 1) * T::SIZE + lane;
            bool have_n = (jn < nrows) && (use32 ? row_range32<D>(dv, uf, x0f, clo, chi, b32, jn, pa_n, pb_n)
                                                 : row_range<D>(dv, q, clo, chi, cen, rho2, jn, pa_n, pb_n));
            HVB_TRACE_EVENT(1, 0);
            HVB_TRACE_EVENT(2, (have && pb > pa) ? pb - pa : 0);
            if (have && pb > pa) scan_points<D>(dv, q, uf, w2f, x0f, flt, pa, pb, st, best, ls);
            have = have_n; pa = pa_n; pb = pb_n;
            // share the best bound and shrink the ball
            if (T::SIZE > 1) {
                best_reduce(tile, best);
#pragma unroll
                for (int m = T::SIZE / 2; m >= 1; m >>= 1) flt.tb2 = fminf(flt.tb2, tile.shfl_xor(flt.tb2, m));
            }
            shrink_ball<D>(dv, q, use32, best, flt, uf, r32, a32, perp2f, Ts32, b32, Ts, cen, rho2, rho, R0, perp2);
        }
        }
        if (T::SIZE > 1) {
#pragma unroll
            for (int m = T::SIZE / 2; m >= 1; m >>= 1) flt.tb2 = fminf(flt.tb2, tile.shfl_xor(flt.tb2, m));
        }
        settle_stage<D>(dv, q, st, flt.tb2, best, ls);
        if (T::SIZE > 1) {
            best_reduce(tile, best);
            st.rejected = tile.ballot(st.rejected) != 0u;
        }
        if (st.rejected) { st.rejected = false; st.tighten = false; continue; }     // redo this stage, FP64 decides everything
        if (best.t <= Ts0 || !(Tst < INFINITY)) break;
        scale *= dv.probe_growth;
    }
    return best;
}

// ------------------------------------------------------------------------------------------------------------
// vertex set + edge table
// ------------------------------------------------------------------------------------------------------------
template <int D>
HVB_HD bool sig_equal(const Dev<D>& dv, u32 v, const int* sig) {
    const int* p = dv.vsig + (size_t)v * (D + 1);
    bool eq = true;
#pragma unroll
    for (int k = 0; k < D + 1; ++k) eq &= (ld_cg(p + k) == sig[k]);
    return eq;
}

// Inserts (sig, r) unless present.  Returns the new index, or 0xffffffff if it was known / the store is full.
// Replaces haskey + push! (abstractmesh.jl:111-153 -> hvdatabase.jl:94-116).
template <int D>
HVB_HD u32 vertex_insert(const Dev<D>& dv, const int* sig, const double* r, LocalStats& ls, u64 h, u64 s) {
    // h = hash of sig, s = already loaded content of its home slot
    u64 fp = (h >> 32) << 32;
    if (fp == 0) fp = 1ULL << 32;
    u64 slot = h & dv.vmask;
    u32 mine = 0xffffffffu;
    for (;;) {
        if (s == 0) {
            if (mine == 0xffffffffu) {
                mine = atom_add(dv.vcount, 1u);
                if (mine >= dv.vcap) { atom_or(&dv.ctr->flags, (u32)FLAG_VFULL); return 0xffffffffu; }
                int* ps = dv.vsig + (size_t)mine * (D + 1);
                double* pr = dv.vr + (size_t)mine * D;
#pragma unroll
                for (int k = 0; k < D + 1; ++k) ps[k] = sig[k];
#pragma unroll
                for (int k = 0; k < D; ++k) pr[k] = r[k];
                mem_fence();
            }
            s = atom_cas(dv.vtab + slot, 0ULL, fp | (u64)(mine + 1u));
            if (s == 0) return mine;
        }
        if ((s >> 32) == (fp >> 32) && sig_equal<D>(dv, (u32)(s & 0xffffffffu) - 1u, sig)) {
            if (mine != 0xffffffffu) { dv.vsig[(size_t)mine * (D + 1)] = -1; ls.dead++; }   // lost a race: dead record
            ls.dup_hits++;
            return 0xffffffffu;
        }
        slot = (slot + 1) & dv.vmask;
        s = ld_cg(dv.vtab + slot);
    }
}

// does the edge stored in slot value s equal (sig minus position k)?
template <int D>
HVB_HD bool edge_equal(const Dev<D>& dv, u64 s, const int* sig, int k) {
    u32 v2 = (u32)((s >> 3) & 0xffffffffULL);
    int k2 = (int)(s & 7);
    const int* p = dv.vsig + (size_t)v2 * (D + 1);
    bool eq = true;
    int i2 = 0;
#pragma unroll
    for (int i = 0; i < D + 1; ++i) {
        if (i == k) continue;
        if (i2 == k2) ++i2;
        eq &= (ld_cg(p + i2) == sig[i]);
        ++i2;
    }
    return eq;
}

// Registers the sub-facet (sig minus position k) of vertex v.  First endpoint: the edge becomes an open frontier
// entry (returns its slot); second endpoint: the edge is closed (returns ~0).  `h` is the edge hash and `s` the
// already loaded content of its home slot.  Replaces pushedge! (edgehashing.jl:66-111) and queue_edges_OnFind
// (edgeiteratebase.jl:128-149).
// the same comparison against a row that is already in registers (row = signature of the vertex named by slot s)
template <int D>
HVB_HD bool edge_equal_row(const int (&row)[D + 1], u64 s, const int* sig, int k) {
    const int k2 = (int)(s & 7);
    bool eq = true;
    int i2 = 0;
#pragma unroll
    for (int i = 0; i < D + 1; ++i) {
        if (i == k) continue;
        if (i2 == k2) ++i2;
        int val = 0;
#pragma unroll
        for (int j = 0; j < D + 1; ++j) val = (j == i2) ? row[j] : val;
        eq &= (val == sig[i]);
        ++i2;
    }
    return eq;
}

template <int D>
HVB_HD u64 edge_register(const Dev<D>& dv, const int* sig, u32 v, int k, u64 h, u64 s, bool have_row, const int (&row)[D + 1]) {
    u64 mine = edge_slot(h, v, k);
    u64 slot = h & dv.emask;
    for (;;) {
        if (s == 0) {
            s = atom_cas(dv.etab + slot, 0ULL, mine);
            if (s == 0) return slot;
        }
        if (((s ^ mine) & EDGE_FPMASK) == 0 && (have_row ? edge_equal_row<D>(row, s, sig, k) : edge_equal<D>(dv, s, sig, k))) {
            if (!(s & EDGE_CLOSED)) atom_or(dv.etab + slot, EDGE_CLOSED);
            return ~0ULL;
        }
        slot = (slot + 1) & dv.emask;
        s = ld_cg(dv.etab + slot);
        have_row = false;
    }
}

// frontier entry: [63:32] edge slot, [31:3] vertex index, [2:0] dropped position
HVB_HD u64 frontier_entry(u64 slot, u32 v, int k) { return (slot << 32) | ((u64)v << 3) | (u64)k; }

// Shared tail of a walk / a descent: store the vertex, register its d+1 sub-facets, append the open ones to the
// next frontier.  sig sorted.  All lanes call; lane 0 inserts, the sub-facets are dealt round-robin to the lanes.
// The home slots of all sub-facets are loaded before any of them is processed (independent loads in flight).
template <int D, class T>
HVB_HD void commit_vertex(const Dev<D>& dv, const T& tile, const int* sig, const double* r,
                          u64* q_out, u32* q_count, u32 q_cap, LocalStats& ls) {
    const int lane = tile.lane();
    // which generators are real and belong to a cell this context explores
    unsigned actmask = 0;
#pragma unroll
    for (int i = 0; i < D + 1; ++i)
        if ((sig[i] < dv.n) && (dv.active[sig[i]] != 0)) actmask |= 1u << i;
    // sub-facet k = pass * G + lane: every lane runs the SAME instructions on its own k (no serialisation by lane)
    const int NPASS = (D + 1 + T::SIZE - 1) / T::SIZE;
    // every independent load of the commit is issued up front: the vertex set's home slot, the home slots of the
    // sub-facets and (d <= 4) the signatures behind matching edge slots
    const u64 hv = hash_ids<D>(sig, D + 1, -1);
    u64 sv = 0;
    if (lane == 0) sv = ld_cg(dv.vtab + (hv & dv.vmask));
    u64 hs[NPASS], s0[NPASS];
    bool mine[NPASS];
#pragma unroll
    for (int ps_ = 0; ps_ < NPASS; ++ps_) {
        int k = ps_ * T::SIZE + lane;
        mine[ps_] = (k < D + 1) && ((actmask & ~(1u << k)) != 0);     // the sub-facet keeps an explored real generator
        hs[ps_] = 0; s0[ps_] = 0;
        if (mine[ps_]) { hs[ps_] = hash_ids<D>(sig, D + 1, k); s0[ps_] = ld_cg(dv.etab + (hs[ps_] & dv.emask)); }
    }
    const bool PREROW = (D <= 4);
    int crow[PREROW ? NPASS : 1][D + 1];
    bool cvalid[NPASS];
#pragma unroll
    for (int ps_ = 0; ps_ < NPASS; ++ps_) {
        cvalid[ps_] = false;
        if (PREROW) {
            cvalid[ps_] = mine[ps_] && s0[ps_] != 0 && (((s0[ps_] ^ edge_slot(hs[ps_], 0u, 0)) & EDGE_FPMASK) == 0);
#pragma unroll
            for (int i = 0; i < D + 1; ++i) crow[PREROW ? ps_ : 0][i] = 0;
            if (cvalid[ps_]) {
                const int* prow = dv.vsig + (size_t)((u32)((s0[ps_] >> 3) & 0xffffffffULL)) * (D + 1);
#pragma unroll
                for (int i = 0; i < D + 1; ++i) crow[PREROW ? ps_ : 0][i] = ld_cg(prow + i);
            }
        }
    }
    u32 v = 0xffffffffu;
    if (lane == 0) v = vertex_insert<D>(dv, sig, r, ls, hv, sv);
    if (T::SIZE > 1) v = tile.shfl(v, 0);
    if (v == 0xffffffffu) return;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < D + 1; ++k)
            if (sig[k] < dv.n) dv.has_vertex[sig[k]] = 1;
    }
#pragma unroll
    for (int ps_ = 0; ps_ < NPASS; ++ps_) {
        if (!mine[ps_]) continue;
        int k = ps_ * T::SIZE + lane;
        u64 slot = edge_register<D>(dv, sig, v, k, hs[ps_], s0[ps_], cvalid[ps_], crow[PREROW ? ps_ : 0]);
        if (slot != ~0ULL) {
            u32 pos = atom_add(q_count, 1u);
            if (pos < q_cap) q_out[pos] = frontier_entry(slot, v, k);
            else atom_or(&dv.ctr->flags, (u32)FLAG_QFULL);
        }
    }
    tile.sync();
}

// insertion of g into the sorted list sig[0..cnt)
HVB_HD void sorted_insert(int* sig, int cnt, int g) {
    int i = cnt;
    while (i > 0 && sig[i - 1] > g) { sig[i] = sig[i - 1]; --i; }
    sig[i] = g;
}

// ------------------------------------------------------------------------------------------------------------
// one frontier entry: walk the open edge from its known endpoint (walkray raycast.jl:125-164 +
// systematic_explore_vertex sysvoronoi.jl:490-525, one edge at a time)
// ------------------------------------------------------------------------------------------------------------
// first half of a walk: loads the origin vertex of frontier entry `item` and builds the ray (direction u_qr
// tools.jl:773-790, half-space threshold raycast.jl:802-804).  false: the direction could not be built.
template <int D>
HVB_HD bool ray_setup(const Dev<D>& dv, u64 item, RayQ<D>& q, int (&sig)[D + 1], u32& v, int& kd) {
    v = (u32)((item >> 3) & 0x1fffffffULL);
    kd = (int)(item & 7);
#pragma unroll
    for (int k = 0; k < D + 1; ++k) { sig[k] = ld_cg(dv.vsig + (size_t)v * (D + 1) + k); q.excl[k] = sig[k]; }
    q.nexcl = D + 1;
#pragma unroll
    for (int k = 0; k < D; ++k) q.r[k] = ld_cg(dv.vr + (size_t)v * D + k);
    // x0 = first generator of the edge (always a real one: plane ids sort last)
    const int id0 = (kd == 0) ? sig[1] : sig[0];
    const int i0 = (kd == 0) ? 1 : 0;
#pragma unroll
    for (int k = 0; k < D; ++k) q.x0[k] = dv.x64[(size_t)id0 * D + k];
    const PlaneSet* ps = dv.planes;
    // direction: orthogonal to the edge's difference vectors / plane normals, away from the dropped generator
    double V[D + 1][D];
    double xd[D];
#pragma unroll
    for (int i = 0; i < D + 1; ++i) {
        int id = sig[i];
        if (id < dv.n) {
#pragma unroll
            for (int k = 0; k < D; ++k) V[i][k] = dv.x64[(size_t)id * D + k] - q.x0[k];
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) V[i][k] = ps->normal[(id - dv.n) * 6 + k];
        }
        if (i == kd) {
#pragma unroll
            for (int k = 0; k < D; ++k) xd[k] = -V[i][k];
        }
    }
    const unsigned emask = ((1u << (D + 1)) - 1u) & ~(1u << kd) & ~(1u << i0);
    if (!ortho_direction<D>(V, emask, xd)) return false;
#pragma unroll
    for (int k = 0; k < D; ++k) q.u[k] = xd[k];
    // half-space threshold: c = max_{g in edge} u.x_g, c += |c| * plane_tol   (raycast.jl:802-804)
    {
        double cm = -INFINITY;
        double ux0 = dotD<D>(q.u, q.x0);
#pragma unroll
        for (int i = 0; i < D + 1; ++i) {
            if (i == kd) continue;
            int id = sig[i];
            double ux;
            if (id < dv.n) ux = dotD<D>(q.u, dv.x64 + (size_t)id * D);
            else {   // mirror image of x0 in the plane
                const double* nrm = ps->normal + (id - dv.n) * 6;
                double nu = 0, nx0 = 0;
#pragma unroll
                for (int k = 0; k < D; ++k) { nu += nrm[k] * q.u[k]; nx0 += nrm[k] * q.x0[k]; }
                ux = ux0 + 2.0 * (ps->off[id - dv.n] - nx0) * nu;
            }
            cm = fmax(cm, ux);
        }
        q.c = cm + fabs(cm) * dv.plane_tol;
    }
    {
        double w[D];
#pragma unroll
        for (int k = 0; k < D; ++k) w[k] = q.x0[k] - q.r[k];
        q.R0sq = dotD<D>(w, w);
        q.a = dotD<D>(q.u, w);
        q.rnorm = sqrt(dotD<D>(q.r, q.r));
    }
    return true;
}

// second half of a walk: the winner of the min-t query becomes an unbounded edge (stored here, returns false) or a
// new vertex (sig2 sorted, r2; returns true)
template <int D>
HVB_HD bool ray_result(const Dev<D>& dv, int lane, const RayQ<D>& q, const int (&sig)[D + 1], u32 v, int kd, const Best& best,
                       int (&sig2)[D + 1], double (&r2)[D], LocalStats& ls) {
    if (best.id < 0) {                       // unbounded edge (sysvoronoi.jl:504-511)
        if (lane == 0) {
            u32 pos = atom_add(dv.ray_count, 1u);
            if (pos < dv.ray_cap) {
                dv.ray_item[pos] = (v << 3) | (u32)kd;
#pragma unroll
                for (int k = 0; k < D; ++k) dv.ray_u[(size_t)pos * D + k] = q.u[k];
            } else atom_or(&dv.ctr->flags, (u32)FLAG_RFULL);
        }
        return false;
    }
    if (near_tie(best, q.R0sq)) { ls.degenerate += (lane == 0); if (lane == 0) atom_or(&dv.ctr->flags, (u32)FLAG_DEGEN); }
    // new vertex: (sig minus position kd) plus the winner, kept sorted with static indexing
    {
        int e[D];
        int pos = 0;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            e[j] = (j < kd) ? sig[j] : sig[j + 1];
            pos += (e[j] < best.id) ? 1 : 0;
        }
#pragma unroll
        for (int i = 0; i < D + 1; ++i) {
            int lo_ = (i < D) ? e[i] : 0;
            int hi_ = (i > 0) ? e[i - 1] : 0;
            sig2[i] = (i < pos) ? lo_ : ((i == pos) ? best.id : hi_);
        }
    }
#pragma unroll
    for (int k = 0; k < D; ++k) r2[k] = q.r[k] + best.t * q.u[k];
    return true;
}

template <int D, class T>
HVB_HD void ray_finish(const Dev<D>& dv, const T& tile, const RayQ<D>& q, const int (&sig)[D + 1], u32 v, int kd, const Best& best,
                       u64* q_out, u32* q_count, u32 q_cap, LocalStats& ls) {
    int sig2[D + 1];
    double r2[D];
    if (ray_result<D>(dv, tile.lane(), q, sig, v, kd, best, sig2, r2, ls)) commit_vertex<D, T>(dv, tile, sig2, r2, q_out, q_count, q_cap, ls);
}

template <int D, class T>
HVB_HD void expand_item(const Dev<D>& dv, const T& tile, u64 item, u64* q_out, u32* q_count, u32 q_cap, LocalStats& ls) {
    // every lane of the tile works on the same snapshot of the entry (the closed bit was checked when it was acquired)
    tile.sync();
    int sig[D + 1];
    RayQ<D> q;
    u32 v; int kd;
    if (!ray_setup<D>(dv, item, q, sig, v, kd)) { ls.seed_fail += (tile.lane() == 0); return; }
    Best best = min_t_query<D, T>(dv, tile, q, ls);
    ray_finish<D, T>(dv, tile, q, sig, v, kd, best, q_out, q_count, q_cap, ls);
}

// ------------------------------------------------------------------------------------------------------------
// seeding: descent (raycast.jl:45-109) from generator `start`
// ------------------------------------------------------------------------------------------------------------
// out-of-line copy for the (cold) descent, which would otherwise inline the query 2 d times
template <int D, class T>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
Best min_t_query_call(const Dev<D>& dv, const T& tile, const RayQ<D>& q, LocalStats& ls) {
    return min_t_query<D, T>(dv, tile, q, ls);
}

HVB_HD double unit_hash(u64& s) {          // deterministic stand-in for randn (raycast.jl:228)
    s = mix64(s + 0x9e3779b97f4a7c15ULL);
    return ((double)(s >> 11) + 0.5) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

template <int D, class T>
HVB_HD void seed_item(const Dev<D>& dv, const T& tile, int start, u64* q_out, u32* q_count, u32 q_cap, LocalStats& ls) {
    const int lane = tile.lane();
    tile.sync();
    ls.seeds += (lane == 0);
    const PlaneSet* ps = dv.planes;
    for (int attempt = 0; attempt < 8; ++attempt) {
        int sig[D + 1];          // in order of discovery; sig[0] = start
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
                    int id = sig[i];
                    if (id < dv.n) {
#pragma unroll
                        for (int k = 0; k < D; ++k) V[i][k] = dv.x64[(size_t)id * D + k] - q.x0[k];
                    } else {
#pragma unroll
                        for (int k = 0; k < D; ++k) V[i][k] = ps->normal[(id - dv.n) * 6 + k];
                    }
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
                double ux0 = dotD<D>(q.u, q.x0);
#pragma unroll
                for (int i = 0; i < D + 1; ++i) {
                    if (i < cnt) {
                        int id = sig[i];
                        double ux;
                        if (id < dv.n) ux = dotD<D>(q.u, dv.x64 + (size_t)id * D);
                        else {
                            const double* nrm = ps->normal + (id - dv.n) * 6;
                            double nu = 0, nx0 = 0;
#pragma unroll
                            for (int k = 0; k < D; ++k) { nu += nrm[k] * q.u[k]; nx0 += nrm[k] * q.x0[k]; }
                            ux = ux0 + 2.0 * (ps->off[id - dv.n] - nx0) * nu;
                        }
                        cm = fmax(cm, ux);
                    }
                }
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
            if (near_tie(best, q.R0sq)) { ls.degenerate += (lane == 0); if (lane == 0) atom_or(&dv.ctr->flags, (u32)FLAG_DEGEN); }
#pragma unroll
            for (int k = 0; k < D; ++k) q.r[k] += best.t * q.u[k];
            sig[cnt++] = best.id;
        }
        if (!ok) continue;
        int ssig[D + 1];
        for (int i = 0; i < D + 1; ++i) sorted_insert(ssig, i, sig[i]);
        commit_vertex<D, T>(dv, tile, ssig, q.r, q_out, q_count, q_cap, ls);
        return;
    }
    ls.seed_fail += (lane == 0);
}

// ------------------------------------------------------------------------------------------------------------
// canonical coordinates: the point equidistant to the generators / on the planes of sig, solved from the
// generators alone so that the output does not depend on the walk that found the vertex (sig is passed in the
// caller's canonical order: ascending ORIGINAL ids).  Replaces
// walkray_correct_vertex / _correct_vertex (raycast.jl:242-318).  Returns the relative variance of the squared
// radii over the real generators (vertex_variance, raycast.jl:320-329).
// ------------------------------------------------------------------------------------------------------------
template <int D>
HVB_HD double canonical_vertex(const Dev<D>& dv, const int* sig, double* r, double* flatness = nullptr) {
    const PlaneSet* ps = dv.planes;
    double x0[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x0[k] = dv.xcan[(size_t)sig[0] * D + k];
    double A[D][D], b[D], z[D];
    double rownorm2 = 1.0;             // product of the squared row norms
#pragma unroll
    for (int i = 0; i < D; ++i) {
        int id = sig[i + 1];
        if (id < dv.n) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < D; ++k) { A[i][k] = dv.xcan[(size_t)id * D + k] - x0[k]; s += A[i][k] * A[i][k]; }
            b[i] = 0.5 * s;
            rownorm2 *= s;
        } else {
            double s = 0;
#pragma unroll
            for (int k = 0; k < D; ++k) { A[i][k] = ps->normal[(id - dv.n) * 6 + k]; s += A[i][k] * x0[k]; }
            b[i] = ps->off[id - dv.n] - s;
        }
    }
    double pivprod = 1.0;              // |det A| = product of the pivots of the first elimination
#pragma unroll
    for (int k = 0; k < D; ++k) z[k] = 0.0;
    // direct solve of A z = b plus one step of iterative refinement (Gaussian elimination with partial pivoting);
    // the walked coordinates are deliberately not used, so the result depends on sig alone
    for (int rep = 0; rep < 2; ++rep) {
        double M[D][D + 1];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double res = b[i];
#pragma unroll
            for (int k = 0; k < D; ++k) { M[i][k] = A[i][k]; res -= A[i][k] * z[k]; }
            M[i][D] = res;
        }
#pragma unroll
        for (int c = 0; c < D; ++c) {
            int pv = c;
            double mx = fabs(M[c][c]);
#pragma unroll
            for (int i = 0; i < D; ++i)
                if (i > c && fabs(M[i][c]) > mx) { mx = fabs(M[i][c]); pv = i; }
#pragma unroll
            for (int i = 0; i < D; ++i)
                if (i == pv && pv != c) {
#pragma unroll
                    for (int k = 0; k < D + 1; ++k) { double tmp = M[c][k]; M[c][k] = M[i][k]; M[i][k] = tmp; }
                }
            double inv = (M[c][c] != 0.0) ? 1.0 / M[c][c] : 0.0;
            if (rep == 0) pivprod *= fabs(M[c][c]);
#pragma unroll
            for (int i = 0; i < D; ++i)
                if (i > c) {
                    double f = M[i][c] * inv;
#pragma unroll
                    for (int k = 0; k < D + 1; ++k)
                        if (k >= c) M[i][k] -= f * M[c][k];
                }
        }
        double dz[D];
#pragma unroll
        for (int c = D - 1; c >= 0; --c) {
            double s = M[c][D];
#pragma unroll
            for (int k = 0; k < D; ++k)
                if (k > c) s -= M[c][k] * dz[k];
            dz[c] = (M[c][c] != 0.0) ? s / M[c][c] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) z[k] += dz[k];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) r[k] = x0[k] + z[k];
    // how far the d + 1 generators are from lying in one hyperplane: |det| of the unit row vectors (1 for an orthogonal corner)
    if (flatness) *flatness = (rownorm2 > 0) ? pivprod / sqrt(rownorm2) : 0.0;
    // variance over the real generators
    double dist[D + 1], mean = 0;
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < D + 1; ++i) {
        dist[i] = 0;
        if (sig[i] < dv.n) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < D; ++k) { double t = dv.xcan[(size_t)sig[i] * D + k] - r[k]; s += t * t; }
            dist[i] = s; mean += s; ++cnt;
        }
    }
    mean /= cnt;
    double var = 0;
#pragma unroll
    for (int i = 0; i < D + 1; ++i)
        if (sig[i] < dv.n) var += (dist[i] - mean) * (dist[i] - mean);
    return var / (mean * mean);
}

}  // namespace hvb
