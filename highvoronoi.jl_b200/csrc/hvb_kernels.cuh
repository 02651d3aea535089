// hvb_kernels.cuh -- __global__ wrappers around hvb_core.cuh plus the index-build and finalize kernels.
// sm_100a only.  Launch shapes: tiles of G lanes per frontier entry, grid-stride over entries, grids sized in
// multiples of the SM count.
#pragma once
#include <cuda_runtime.h>
#include "hvb_core.cuh"
#include "hvb_geometry.cuh"
#include "hvb_hull.cuh"
#include "hvb_wrap.cuh"
#include "hvb_nongeneral.hpp"

namespace hvb {

// ------------------------------------------------------------------------------------------------------------
// statistics: per-thread registers -> one atomic per warp at kernel end
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void flush_stats(const LocalStats& ls, Counters* c) {
    const u32* src = reinterpret_cast<const u32*>(&ls);
    u64* dst = reinterpret_cast<u64*>(c);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(LocalStats) / sizeof(u32)); ++i) {
        u32 v = src[i];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst + i, (u64)v);
    }
}

// ------------------------------------------------------------------------------------------------------------
// spatial index: counting sort of the generators into grid cells (replaces the KD-tree build, kd_tree.jl:27-158)
// ------------------------------------------------------------------------------------------------------------
// bounding box of the generators + domain check (check_boundary, boundary.jl:437): per-block partials, reduced by
// the last block to finish.  out: [2*D] doubles (min, max), viol: smallest index of a generator outside a plane
// (or non-finite), 0xffffffff if none
template <int D>
static __global__ void k_bbox_check(const double* __restrict__ xs, int n, const PlaneSet* __restrict__ ps, double* __restrict__ partial,
                             unsigned int* __restrict__ done, double* __restrict__ out, unsigned int* __restrict__ viol) {
    double mn[D], mx[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { mn[k] = 1e300; mx[k] = -1e300; }
    unsigned int bad = 0xffffffffu;
    const int P = ps->P;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double x[D];
        bool ok = true;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            x[k] = xs[(size_t)i * D + k];
            ok &= (x[k] == x[k]) && (fabs(x[k]) <= 1e150);
            mn[k] = fmin(mn[k], x[k]); mx[k] = fmax(mx[k], x[k]);
        }
        for (int p = 0; p < P; ++p) {
            double sdot = 0;
#pragma unroll
            for (int k = 0; k < D; ++k) sdot += ps->normal[p * 6 + k] * x[k];
            ok &= !(sdot > ps->off[p]);
        }
        if (!ok) bad = min(bad, (unsigned int)i);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], m));
            mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], m));
        }
        bad = min(bad, __shfl_xor_sync(0xffffffffu, bad, m));
    }
    __shared__ double sm[8][2 * D];
    __shared__ unsigned int sbad[8];
    __shared__ bool last;
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) { sm[w][k] = mn[k]; sm[w][D + k] = mx[k]; }
        sbad[w] = bad;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int j = 1; j < nw; ++j) {
#pragma unroll
            for (int k = 0; k < D; ++k) { sm[0][k] = fmin(sm[0][k], sm[j][k]); sm[0][D + k] = fmax(sm[0][D + k], sm[j][D + k]); }
            sbad[0] = min(sbad[0], sbad[j]);
        }
#pragma unroll
        for (int k = 0; k < 2 * D; ++k) partial[(size_t)blockIdx.x * 2 * D + k] = sm[0][k];
        if (sbad[0] != 0xffffffffu) atomicMin(viol, sbad[0]);
        __threadfence();
        last = (atomicAdd(done, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        // the last block folds the per-block partials with all its threads (a serial loop of one thread over ~1000 partials
        // was the longest part of this kernel)
        double r[2 * D];
#pragma unroll
        for (int k = 0; k < D; ++k) { r[k] = 1e300; r[D + k] = -1e300; }
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                r[k] = fmin(r[k], __ldcg(partial + (size_t)b * 2 * D + k));
                r[D + k] = fmax(r[D + k], __ldcg(partial + (size_t)b * 2 * D + D + k));
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                r[k] = fmin(r[k], __shfl_xor_sync(0xffffffffu, r[k], m));
                r[D + k] = fmax(r[D + k], __shfl_xor_sync(0xffffffffu, r[D + k], m));
            }
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int k = 0; k < 2 * D; ++k) sm[w][k] = r[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int nw = blockDim.x >> 5;
            for (int j = 1; j < nw; ++j) {
#pragma unroll
                for (int k = 0; k < D; ++k) { sm[0][k] = fmin(sm[0][k], sm[j][k]); sm[0][D + k] = fmax(sm[0][D + k], sm[j][D + k]); }
            }
#pragma unroll
            for (int k = 0; k < 2 * D; ++k) out[k] = sm[0][k];
        }
    }
}

template <int D>
static __global__ void k_cell_count(Dev<D> dv, const double* __restrict__ xs, int* __restrict__ cell_of, int* __restrict__ cell_cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dv.n) return;
    double x[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = xs[(size_t)i * D + k];
    int c = cell_index<D>(dv, x);
    cell_of[i] = c;
    atomicAdd(cell_cnt + c, 1);
}

template <int D>
static __global__ void k_scatter(Dev<D> dv, const double* __restrict__ xs, const int* __restrict__ cell_of,
                          const int* __restrict__ cell_start, int* __restrict__ cursor,
                          double* __restrict__ x64, float* __restrict__ x32, int* __restrict__ perm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dv.n) return;
    int c = cell_of[i];
    int pos = cell_start[c] + atomicAdd(cursor + c, 1);
    perm[pos] = i;                    // coordinates are written by k_cell_sort once the order inside the cell is fixed
}

// deterministic order inside a cell: sort each cell's entries by caller id (cells hold a handful of points)
static __global__ void k_cell_sort(const int* __restrict__ cell_start, int ncells, int* __restrict__ perm) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    int a = cell_start[c], b = cell_start[c + 1];
    for (int i = a + 1; i < b; ++i) {
        int key = perm[i];
        int j = i - 1;
        while (j >= a && perm[j] > key) { perm[j + 1] = perm[j]; --j; }
        perm[j + 1] = key;
    }
}
// coordinates in grid order (FP64 for verification, FP32 minus the grid origin for the filter) and the inverse
// permutation; one thread per sorted position: the writes are coalesced
// Non-general position (Ctx::resolve_degenerate): the generators are moved by a deterministic pseudo-random offset of relative
// size `rel` (times the extent of the cloud) -- an explicit simulation of simplicity.  The offset depends on the caller's id
// and the axis alone.
static __global__ void k_perturb(const double* __restrict__ src, double* __restrict__ dst, size_t count, int dim, double amp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    dst[i] = src[i] + amp * perturb_unit((u64)i);
}
// caller coordinates in grid order (the coordinates canonical_vertex solves from when the search ran on perturbed ones)
template <int D>
static __global__ void k_gather_canon(const double* __restrict__ xs, const int* __restrict__ perm, int n, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int o = perm[i];
#pragma unroll
    for (int k = 0; k < D; ++k) out[(size_t)i * D + k] = xs[(size_t)o * D + k];
}
template <int D>
static __global__ void k_gather_points(Dev<D> dv, const double* __restrict__ xs, const int* __restrict__ perm,
                                double* __restrict__ x64, float* __restrict__ x32, int* __restrict__ inv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dv.n) return;
    const int o = perm[i];
    inv[o] = i;
#pragma unroll
    for (int k = 0; k < X32<D>::STRIDE; ++k) {
        double v = (k < D) ? xs[(size_t)o * D + (k < D ? k : 0)] : 0.0;
        if (k < D) x64[(size_t)i * D + k] = v;
        x32[(size_t)i * X32<D>::STRIDE + k] = (k < D) ? (float)(v - dv.lo[k < D ? k : 0]) : 0.f;
    }
}

// active[g] = 1 for the sorted positions of this context's slab (parallelmesh.jl:52-87) or of the Iter cells
// ------------------------------------------------------------------------------------------------------------
// multi-GPU decomposition: owner[p] = rank that explores the cell of sorted position p.
//   slabs  (decomposition 0): contiguous ranges of the sorted order with equal counts -- partition_indices,
//           parallelmesh.jl:52-87.  The order is cell-linear with axis 0 slowest, so these are slabs across axis 0.
//   blocks (decomposition 1): the domain is cut along up to three axes (m[0] x m[1] x m[2] = world) at the quantiles of the
//           points' coordinates.  A rank finds every vertex that touches its cells, so the layer of vertices two
//           neighbouring ranks both find grows with the SURFACE of a part: 8 slabs have 14 faces of full cross-section,
//           2 x 2 x 2 blocks have 6 -- and all blocks of a cube are corner blocks, alike in their share of cheap boundary
//           cells (with slabs the two end ranks of C4 had 2/3 of the work of the inner ones).
// ------------------------------------------------------------------------------------------------------------
struct BlockSpec {
    int mode;                 // 0 slabs, 1 blocks
    int world;
    int m[3];                 // parts along axes 0, 1, 2 (1 = axis not cut)
    double cut[3][17];        // blocks: coordinate where part j of axis a begins (j = 1 .. m - 1): a point with x_a >= cut belongs
                              // to part j or a later one.  Quantiles of the points' coordinates (1024-bin histogram), NOT cell
                              // boundaries: a grid of high dimension has ~10 cells per axis, far too coarse to halve evenly
    long long bound[65];      // slabs: sorted position where slab k begins (bound[world] = n)
};
// histogram of the points' coordinates along axis a (HVB_CUT_BINS bins over [lo, lo + width)), per-block in shared memory
#define HVB_CUT_BINS 1024
template <int D>
static __global__ void k_coord_hist(const double* __restrict__ x64, int n, int a, double lo, double inv_w, unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[HVB_CUT_BINS];
    for (int i = threadIdx.x; i < HVB_CUT_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        int b = (int)((x64[(size_t)p * D + a] - lo) * inv_w * HVB_CUT_BINS);
        b = b < 0 ? 0 : (b >= HVB_CUT_BINS ? HVB_CUT_BINS - 1 : b);
        atomicAdd(sh + b, 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HVB_CUT_BINS; i += blockDim.x) if (sh[i]) atomicAdd(hist + i, sh[i]);
}
template <int D>
static __global__ void k_assign_owner(Dev<D> dv, BlockSpec bs, unsigned char* __restrict__ owner) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dv.n) return;
    int r = 0;
    if (bs.mode == 0) {
        while (r + 1 < bs.world && (long long)p >= bs.bound[r + 1]) ++r;
    } else {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (a >= D) break;
            const double x = dv.x64[(size_t)p * D + a];
            int b = 0;
            while (b + 1 < bs.m[a] && x >= bs.cut[a][b + 1]) ++b;
            r = r * bs.m[a] + b;
        }
    }
    owner[p] = (unsigned char)r;
}
// active[p] = 1 for the sorted positions this rank explores (periodic contexts: the caller's own generators only)
static __global__ void k_fill_active_owner(unsigned char* active, const unsigned char* __restrict__ owner, const int* __restrict__ perm,
                                    int n, int n_user, int rank) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) active[i] = ((owner ? owner[i] == rank : true) && perm[i] < n_user) ? 1 : 0;
}
static __global__ void k_inverse_perm(const int* __restrict__ perm, int* __restrict__ inv, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[perm[i]] = i;
}
static __global__ void k_mark_cells(const long long* __restrict__ cells, long long ncells, const int* __restrict__ inv,
                             unsigned char* active, int n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < ncells) { long long c = cells[i] - 1; if (c >= 0 && c < n) active[inv[c]] = 1; }
}

// ------------------------------------------------------------------------------------------------------------
// periodic domains: halo generators (reflect_nodes domain.jl:338-390, iteratively_reflected_points! :310-336) and the
// certificate that replaces the reference's repeat-until-stable periodize! passes (domain.jl:139-166)
// ------------------------------------------------------------------------------------------------------------
// Periodic plane pairs: plane a and its partner b have opposite normals; T = n_a * width moves a point one period
// in direction n_a.  A halo copy is x + sum_p k_p T_p with integer multiplicities |k_p| <= K_p, not all zero, kept
// if it lies inside the periodic planes pushed outwards by the margin (the PlaneSet on the device holds the pushed
// planes: expand_internal_boundary, domain.jl:145).
#define HVB_MAX_PAIRS 8
struct HaloSpec {
    int npairs;
    int K[HVB_MAX_PAIRS];
    int plane_a[HVB_MAX_PAIRS], plane_b[HVB_MAX_PAIRS];
    double T[HVB_MAX_PAIRS][6];
    int ncodes;
};

// one thread per caller generator, all shift codes in ascending order: deterministic halo numbering without atomics.
// Pass 1 (xs_out == nullptr) counts the accepted copies of every generator, pass 2 writes them behind offs[i].
template <int D>
static __global__ void k_halo(const double* xs, int n_user, HaloSpec hs, const PlaneSet* __restrict__ ps,
                       int* __restrict__ counts, const int* __restrict__ offs, double* xs_out,
                       int* __restrict__ origin, signed char* __restrict__ mult) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_user) return;
    double x[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = xs[(size_t)i * D + k];
    int cnt = 0;
    const int base = xs_out ? offs[i] : 0;
    for (int code = 0; code < hs.ncodes; ++code) {
        int km[HVB_MAX_PAIRS];
        int rem = code;
        bool zero = true;
        for (int p = 0; p < hs.npairs; ++p) {
            int w = 2 * hs.K[p] + 1;
            km[p] = rem % w - hs.K[p];
            rem /= w;
            zero &= (km[p] == 0);
        }
        if (zero) continue;
        double y[D];
#pragma unroll
        for (int k = 0; k < D; ++k) y[k] = x[k];
        for (int p = 0; p < hs.npairs; ++p) {
#pragma unroll
            for (int k = 0; k < D; ++k) y[k] += (double)km[p] * hs.T[p][k];
        }
        bool ok = true;
        for (int p = 0; p < hs.npairs && ok; ++p) {
            // the same expression as the domain check (k_bbox_check): a kept copy passes it bit for bit
            double sa = 0, sb = 0;
#pragma unroll
            for (int k = 0; k < D; ++k) { sa += ps->normal[hs.plane_a[p] * 6 + k] * y[k]; sb += ps->normal[hs.plane_b[p] * 6 + k] * y[k]; }
            ok = !(sa > ps->off[hs.plane_a[p]]) && !(sb > ps->off[hs.plane_b[p]]);
        }
        if (!ok) continue;
        if (xs_out) {
            int pos = base + cnt;
#pragma unroll
            for (int k = 0; k < D; ++k) xs_out[(size_t)(n_user + pos) * D + k] = y[k];
            origin[pos] = i;
            for (int p = 0; p < hs.npairs; ++p) mult[(size_t)pos * hs.npairs + p] = (signed char)km[p];
        }
        ++cnt;
    }
    if (!xs_out) counts[i] = cnt;
}



// Certificate of a periodic result.  A vertex that touches one of the caller's generators is a vertex of the
// periodic tessellation if its (empty) ball stays inside the pushed periodic planes: then no periodic image that was
// NOT copied can lie in the ball.  If this holds for all vertices and none of them sits on a pushed plane, every
// cell of a caller generator is exactly its periodic cell.  Per row: the excess of the ball
// over every ORIGINAL periodic plane (-> smallest sufficient margin), whether it touches a pushed periodic plane,
// and the flag "canonical representative": among the periodic images of a vertex that touch caller generators
// exactly one has its smallest-origin generator unshifted.
struct PeriodicCert {
    int nplanes;
    int is_periodic[HVB_MAX_PLANES];
    double off_orig[HVB_MAX_PLANES];
};
struct CertOut {
    unsigned long long max_excess_bits;   // max over rows of (n_p . r + R - off_orig_p), >= 0, as double bits
    unsigned int on_pushed_plane;         // rows with a pushed periodic plane in their signature
    unsigned int canonical;               // rows flagged canonical
    unsigned int self_neighbor;           // rows holding two images of the same generator (domain too small)
    unsigned int pad;
};
template <int D>
static __global__ void k_certify(const long long* __restrict__ sig, const double* __restrict__ r, u32 nv, long long n_user, long long n_ext,
                          const double* __restrict__ xs_ext, const int* __restrict__ halo_origin, const PlaneSet* __restrict__ ps,
                          PeriodicCert pc, unsigned char* __restrict__ vflags, CertOut* __restrict__ out) {
    u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    long long s[D + 1];
#pragma unroll
    for (int k = 0; k < D + 1; ++k) s[k] = sig[(size_t)v * (D + 1) + k];
    double c[D];
#pragma unroll
    for (int k = 0; k < D; ++k) c[k] = r[(size_t)v * D + k];
    // radius: distance to the first generator (rows are sorted, planes last; s[0] is always a generator)
    double R2 = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) { double t = xs_ext[(size_t)(s[0] - 1) * D + k] - c[k]; R2 += t * t; }
    const double R = sqrt(R2);
    bool pushed = false, self_nb = false;
    long long min_origin = 0x7fffffffffffffffLL;
    bool min_unshifted = false;
#pragma unroll
    for (int k = 0; k < D + 1; ++k) {
        if (s[k] > n_ext) { pushed |= (pc.is_periodic[(int)(s[k] - n_ext - 1)] != 0); continue; }
        long long o = (s[k] <= n_user) ? s[k] - 1 : (long long)halo_origin[s[k] - n_user - 1];
        if (o < min_origin) { min_origin = o; min_unshifted = (s[k] <= n_user); }
        else if (o == min_origin && s[k] <= n_user) min_unshifted = true;
    }
    // images of the same generator further down the row (not only of the smallest origin)
#pragma unroll
    for (int a = 0; a < D + 1; ++a)
#pragma unroll
        for (int b = a + 1; b < D + 1; ++b)
            if (s[a] <= n_ext && s[b] <= n_ext) {
                long long oa = (s[a] <= n_user) ? s[a] - 1 : (long long)halo_origin[s[a] - n_user - 1];
                long long ob = (s[b] <= n_user) ? s[b] - 1 : (long long)halo_origin[s[b] - n_user - 1];
                self_nb |= (oa == ob);
            }
    double excess = 0;
    for (int p = 0; p < pc.nplanes; ++p) {
        if (!pc.is_periodic[p]) continue;
        double sdot = 0;
#pragma unroll
        for (int k = 0; k < D; ++k) sdot += ps->normal[p * 6 + k] * c[k];
        excess = fmax(excess, sdot + R - pc.off_orig[p]);
    }
    vflags[v] = min_unshifted ? 1 : 0;
    if (excess > 0) atomicMax(&out->max_excess_bits, (unsigned long long)__double_as_longlong(excess));
    if (pushed) atomicAdd(&out->on_pushed_plane, 1u);
    if (min_unshifted) atomicAdd(&out->canonical, 1u);
    if (self_nb) atomicAdd(&out->self_neighbor, 1u);
}

// ------------------------------------------------------------------------------------------------------------
// seeding and frontier rounds
// ------------------------------------------------------------------------------------------------------------
template <int D, int G>
static __global__ void __launch_bounds__(128) k_seed(Dev<D> dv, const int* __restrict__ seeds, int nseeds, int stride,
                                              u64* q_out, u32* q_count, u32 q_cap) {
    TileDev<G> tile;
    LocalStats ls = {};
    const int tiles_per_block = blockDim.x / G;
    const int ntiles = gridDim.x * tiles_per_block;
    for (int it = blockIdx.x * tiles_per_block + threadIdx.x / G; it < nseeds; it += ntiles) {
        int start = seeds ? seeds[it] : it * stride;
        if (start < dv.n && dv.active[start]) seed_item<D, TileDev<G> >(dv, tile, start, q_out, q_count, q_cap, ls);
    }
    flush_stats(ls, dv.ctr);
}

// resident blocks per SM the walk kernel is compiled for: 4 (<= 128 registers per thread) measured best for d = 2..5
// although the higher dimensions spill: occupancy matters more to this latency-bound kernel
// (profiles/r1_tile_occupancy_sweep.md)
// Vertices the caller's mesh already holds (refinement callers, meshrefine.jl:199-215): converted to internal
// numbering, stored and registered like found vertices, so that the walk continues from them and never returns them.
template <int D>
static __global__ void k_insert_seeds(Dev<D> dv, const long long* __restrict__ sig_in, const double* __restrict__ r_in, long long nseed,
                               int stride, const int* __restrict__ inv, u64* q_out, u32* q_count, u32 q_cap, u32* bad) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nseed) return;
    LocalStats ls = {};
    TileDev<1> tile;
    int sig[D + 1];
    int cnt = 0;
    bool ok = true;
    for (int k = 0; k < stride; ++k) {
        long long g = sig_in[i * stride + k];
        if (g == 0) continue;                                  // unused entry
        if (cnt == D + 1) { ok = false; break; }
        int id;
        if (g >= 1 && g <= dv.n) id = inv[g - 1];
        else if (g > dv.n && g <= dv.n + dv.planes->P) id = (int)(g - 1);
        else { ok = false; break; }
        int j = cnt++;
        while (j > 0 && sig[j - 1] > id) { sig[j] = sig[j - 1]; --j; }
        sig[j] = id;
    }
    if (!ok || cnt != D + 1) { atomicAdd(bad, 1u); return; }   // only general vertices (dim+1 generators) are accepted
    for (int k = 0; k < D; ++k) ok &= (sig[k] != sig[k + 1]);
    if (!ok || sig[0] >= dv.n) { atomicAdd(bad, 1u); return; }
    double r[D];
#pragma unroll
    for (int k = 0; k < D; ++k) r[k] = r_in[i * D + k];
    commit_vertex<D, TileDev<1> >(dv, tile, sig, r, q_out, q_count, q_cap, ls);
}

#ifndef HVB_EXPAND_MINB
#define HVB_EXPAND_MINB 4
#endif
// One frontier round.  Tiles pull entries from a shared cursor and skip closed edges while acquiring, so that all
// tiles of a warp enter the expensive part (direction, min-t query, commit) with live work.
template <int D, int G>
static __global__ void __launch_bounds__(128, HVB_EXPAND_MINB) k_expand(Dev<D> dv, const u64* __restrict__ q_in, const u32* __restrict__ n_in_ptr,
                                                                 u32* cursor, u64* q_out, u32* q_count, u32 q_cap) {
    TileDev<G> tile;
    LocalStats ls = {};
    const u32 n_in = min(*n_in_ptr, q_cap);
    for (;;) {
        u64 item = 0;
        int live = 0;
        if (tile.lane() == 0) {
            for (;;) {
                u32 idx = atomicAdd(cursor, 1u);
                if (idx >= n_in) break;
                u64 it = q_in[idx];
                u64 s = __ldcg(dv.etab + (u32)(it >> 32));
                if (s & EDGE_CLOSED) { ls.closed_skips++; continue; }
                item = it; live = 1;
                break;
            }
        }
        if (G > 1) { item = tile.shfl(item, 0); live = tile.shfl(live, 0); }
        // warp-uniform loop: every lane stays until no tile of the warp has work, and all tiles start their entries
        // together (without this the lanes drift apart for good and the warp executes one lane at a time)
        if (!__any_sync(0xffffffffu, live)) break;
        if (live) expand_item<D, TileDev<G> >(dv, tile, item, q_out, q_count, q_cap, ls);
        __syncwarp();
    }
    flush_stats(ls, dv.ctr);
}

// The whole frontier walk as ONE launch: a single append-only queue (every edge enters it exactly once), consumers
// take tickets (atomicAdd on `head`) and wait for the entry behind their ticket to be published, producers append at
// `tail`.  No round barrier: an edge opened by a new vertex is walked as soon as a lane is free, so there are no
// per-round tails and no host round trips.  Termination: `done` counts fully processed entries (including their
// pushes); when done == tail and a lane's ticket is >= tail nothing can ever be published again.
// A lane never blocks: if its entry is not there yet it idles for one trip of the warp-uniform loop.
struct WalkQueue {
    u64* q;            // [cap], 0xff..ff = not yet published
    u32* tail;         // entries appended so far (also q_count of commit_vertex)
    u32* head;         // tickets handed out
    u32* done;         // entries fully processed
    u32* abort;        // set when the safety timeout fires
    u32 cap;
    u32 stop_on_degenerate;
};
#define HVB_Q_EMPTY 0xffffffffffffffffULL

template <int D, int G>
static __global__ void __launch_bounds__(128, HVB_EXPAND_MINB) k_walk(Dev<D> dv, WalkQueue wq) {
    TileDev<G> tile;
    LocalStats ls = {};
    u32 ticket = 0xffffffffu;          // held by lane 0 of the tile
    u32 my_done = 0;                   // processed entries not yet added to *done
    long long t_idle = clock64();
    for (u32 trip = 0;; ++trip) {
        u64 item = 0;
        int live = 0, finished = 0;
        if (tile.lane() == 0) {
            // non-general position stops the walk at once (the host reports HVB_EDEGENERATE); so does the safety abort
            // (so does a full vertex store / queue: the host grows the tables and restarts)
            {
                const u32 fl = __ldcg(&dv.ctr->flags);
                if ((fl & FLAG_OVERFLOW_MASK) || (wq.stop_on_degenerate && (fl & FLAG_DEGEN)) || __ldcg(wq.abort)) finished = 1;
            }
            for (int tries = 0; tries < 4 && !live && !finished; ++tries) {
                if (ticket == 0xffffffffu) ticket = atomicAdd(wq.head, 1u);
                // a ticket beyond the queue's capacity can never be served (pushes beyond it are refused and flagged
                // by commit_vertex): such a lane only waits for the end like one whose entry is not published yet
                u64 it = (ticket < wq.cap) ? __ldcg(wq.q + ticket) : HVB_Q_EMPTY;
                if (it == HVB_Q_EMPTY) {
                    // nothing behind this ticket yet: publish my progress, then test for global completion
                    if (my_done) { __threadfence(); atomicAdd(wq.done, my_done); my_done = 0; }
                    u32 dn = __ldcg(wq.done);
                    u32 tl = __ldcg(wq.tail);
                    if ((dn == tl && ticket >= tl) || __ldcg(wq.abort)) finished = 1;
                    break;
                }
                u64 s = __ldcg(dv.etab + (u32)(it >> 32));
                if (!(s >> 63)) break;                          // the edge slot is not visible yet: retry next trip
                ticket = 0xffffffffu;
                if (s & EDGE_CLOSED) { ls.closed_skips++; ++my_done; continue; }
                item = it; live = 1;
            }
        }
        if (G > 1) { item = tile.shfl(item, 0); live = tile.shfl(live, 0); finished = tile.shfl(finished, 0); }
        if (__all_sync(0xffffffffu, finished)) break;
        const bool any_live = __any_sync(0xffffffffu, live) != 0;      // voted by all lanes BEFORE they part ways
        if (live) {
            expand_item<D, TileDev<G> >(dv, tile, item, wq.q, wq.tail, wq.cap, ls);
            if (tile.lane() == 0) ++my_done;
        }
        if (any_live) t_idle = clock64();
        else {
            // the whole warp is idle: back off, and make sure a bug can never hang the device (20 s without work)
            __nanosleep(256);
            if ((trip & 1023u) == 1023u && clock64() - t_idle > 40000000000LL) atomicExch(wq.abort, 1u);
        }
        __syncwarp();
    }
    if (tile.lane() == 0 && my_done) { __threadfence(); atomicAdd(wq.done, my_done); }
    flush_stats(ls, dv.ctr);
}

// cells of this context that still have no vertex (sysvoronoi.jl:416-429: they get their own descent)
template <int D>
static __global__ void k_unseeded(Dev<D> dv, int* list, u32* count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dv.n) return;
    if (dv.active[i] && !dv.has_vertex[i]) list[atomicAdd(count, 1u)] = i;
}


// ------------------------------------------------------------------------------------------------------------
// convex hull by the facet walk (hvb_hull.cuh; replaces systematic_chull, chull.jl:241-387)
// ------------------------------------------------------------------------------------------------------------
// generator with the largest coordinate along `axis` (search_max, chull.jl:244): packed {ordered coordinate bits, position}
template <int D>
static __global__ void k_argmax_axis(const double* __restrict__ x64, int n, int axis, unsigned long long* __restrict__ out) {
    unsigned long long best = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float v = (float)x64[(size_t)i * D + axis];
        unsigned int b = __float_as_uint(v);
        b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);           // order-preserving map of float bits
        const unsigned long long key = ((unsigned long long)b << 32) | (unsigned int)i;
        best = key > best ? key : best;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, m); best = o > best ? o : best; }
    if ((threadIdx.x & 31) == 0) atomicMax(out, best);
}
template <int D>
static __global__ void k_hull_seed(Dev<D> dv, HullDev<D> hd, int start, int axis, u64* q_out, u32* q_count, u32 q_cap) {
    TileDev<32> tile;                    // one warp: the descent and the climb are a chain of queries, each shared by the lanes
    LocalStats ls = {};
    hull_seed<D, TileDev<32> >(dv, hd, tile, start, axis, q_out, q_count, q_cap, ls);
    __syncwarp();
    flush_stats(ls, dv.ctr);
}
// one round of walks: a warp per entry (the query of an unbounded edge scans a half-space of the grid), entries pulled from
// a shared cursor
template <int D>
static __global__ void __launch_bounds__(128, HVB_EXPAND_MINB) k_hull_expand(Dev<D> dv, HullDev<D> hd, const u64* __restrict__ q_in, const u32* __restrict__ n_in_ptr,
                                                                      u32* cursor, u64* q_out, u32* q_count, u32 q_cap) {
    TileDev<32> tile;
    LocalStats ls = {};
    const u32 n_in = min(*n_in_ptr, q_cap);
    for (;;) {
        u32 idx = 0;
        if (tile.lane() == 0) idx = atomicAdd(cursor, 1u);
        idx = tile.shfl(idx, 0);
        if (idx >= n_in) break;
        hull_step<D, TileDev<32> >(dv, hd, tile, q_in[idx], q_out, q_count, q_cap, ls);
        __syncwarp();
    }
    flush_stats(ls, dv.ctr);
}
// facets in caller numbering, into the arrays of the unbounded edges (a hull facet IS an unbounded edge): edge = the d
// generators (sorted, 1-based), base = a point the edge starts at, dir = outward unit normal, node = smallest generator
template <int D>
static __global__ void k_final_facets(Dev<D> dv, HullDev<D> hd, const int* __restrict__ perm, u32 nf,
                               long long* __restrict__ edge, double* __restrict__ base, double* __restrict__ dir, long long* __restrict__ node,
                               u32* __restrict__ out_count) {
    u32 f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int* fs = hd.fsig + (size_t)f * D;
    if (fs[0] < 0) return;
    const u32 o = atomicAdd(out_count, 1u);
    long long e[D];
#pragma unroll
    for (int k = 0; k < D; ++k) e[k] = (long long)perm[fs[k]] + 1;
    for (int a = 1; a < D; ++a) { long long key = e[a]; int b = a - 1; while (b >= 0 && e[b] > key) { e[b + 1] = e[b]; --b; } e[b + 1] = key; }
    const u32 v = hd.fitem[f] >> 3;
    for (int k = 0; k < D; ++k) {
        edge[(size_t)o * D + k] = e[k];
        base[(size_t)o * D + k] = dv.vr[(size_t)v * D + k];
        dir[(size_t)o * D + k] = hd.fu[(size_t)f * D + k];
    }
    node[o] = e[0];
}

// ------------------------------------------------------------------------------------------------------------
// convex hull by gift wrapping (hvb_wrap.cuh): three kernels per round, all sized by device-side counts so that the host
// enqueues rounds back to back and looks at the queue only every few rounds
// ------------------------------------------------------------------------------------------------------------
// extreme generators along every axis (search_max, chull.jl:244), exact in FP64: pass 0 the extreme values (order-preserving
// map of the double's bits; slot 2 k: largest along axis k, slot 2 k + 1: smallest), pass 1 the smallest position that
// holds each
template <int D>
static __global__ void k_wrap_extremes(const double* __restrict__ x64, int n, unsigned long long* __restrict__ best_val, int pass,
                                       unsigned int* __restrict__ best_pos) {
    unsigned long long best[2 * D];
    unsigned int pos[2 * D];
#pragma unroll
    for (int a = 0; a < 2 * D; ++a) { best[a] = pass ? __ldcg(best_val + a) : 0ULL; pos[a] = 0xffffffffu; }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            unsigned long long b = (unsigned long long)__double_as_longlong(x64[(size_t)i * D + k]);
            b = (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
            const unsigned long long nb = ~b;
            if (!pass) { best[2 * k] = b > best[2 * k] ? b : best[2 * k]; best[2 * k + 1] = nb > best[2 * k + 1] ? nb : best[2 * k + 1]; }
            else {
                if (b == best[2 * k]) pos[2 * k] = min(pos[2 * k], (unsigned int)i);
                if (nb == best[2 * k + 1]) pos[2 * k + 1] = min(pos[2 * k + 1], (unsigned int)i);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 2 * D; ++a) {
        if (!pass) {
            unsigned long long v = best[a];
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m); v = o > v ? o : v; }
            if ((threadIdx.x & 31) == 0) atomicMax(best_val + a, v);
        } else {
            unsigned int v = pos[a];
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, m));
            if ((threadIdx.x & 31) == 0 && v != 0xffffffffu) atomicMin(best_pos + a, v);
        }
    }
}
// the first seed steps: rotate the hyperplanes {x_k = max} and {x_k = min} about the extreme generators
template <int D>
static __global__ void k_wrap_init(WrapDev<D> wd, const unsigned int* __restrict__ start) {
    for (int a = 0; a < 2 * D; ++a) {
        WrapSeed& sd = wd.seed[a];
        sd.ids[0] = (int)start[a]; sd.cnt = 1;
        for (int k = 0; k < D; ++k) sd.u[k] = (k == (a >> 1)) ? ((a & 1) ? -1.0 : 1.0) : 0.0;
        wd.q[0][a] = WRAP_SEED_ENTRY + (u64)a;
    }
    wd.qcount[0] = 2 * D; wd.qcount[1] = 0; wd.nq[0] = 0; wd.nq[1] = 0;
}
template <int D>
static __global__ void __launch_bounds__(128) k_wrap_prepare(Dev<D> dv, HullDev<D> hd, WrapDev<D> wd, int cur) {
    LocalStats ls = {};
    const u32 cnt = min(__ldcg(wd.qcount + cur), wd.qcap);
    if (blockIdx.x == 0 && threadIdx.x == 0) { wd.qcount[1 - cur] = 0; wd.nq[1 - cur] = 0; }
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        WrapQuery<D> w;
        if (!wrap_prepare<D>(dv, hd, wd, wd.q[cur][i], w, ls)) continue;
        const u32 qi = atomicAdd(wd.nq + cur, 1u);
        if (qi < wd.wq_cap) wd.wq[qi] = w;
        else atomicOr(&dv.ctr->flags, (u32)FLAG_QFULL);
    }
    __syncwarp();
    flush_stats(ls, dv.ctr);
}

__device__ __forceinline__ u32 wrap_smem(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wrap_mbar_init(void* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wrap_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void wrap_mbar_expect(void* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wrap_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wrap_bulk_load(void* dst, const void* src, u32 bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(wrap_smem(dst)), "l"(src), "r"(bytes), "r"(wrap_smem(bar)) : "memory");
}
__device__ __forceinline__ void wrap_mbar_wait(void* bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WRAP_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WRAP_DONE;\n"
        "bra WRAP_WAIT;\n"
        "WRAP_DONE:\n"
        "}\n" ::"r"(wrap_smem(bar)), "r"(parity) : "memory");
}

// The stream.  A block of 4 warps takes work items (query tile, chunk of the generators): QW of its warps own 32 queries
// each (one per lane), the other PW = 4 / QW split every staged tile among them.  The FP32 coordinates of the chunk travel
// through a ring of HVB_WRAP_NS shared-memory tiles filled by the TMA unit (one elected thread issues cp.async.bulk, the
// bytes arrive on the tile's mbarrier); all lanes of a warp read the same staged generator, so a tile is read once per
// warp and broadcast.  Partial results (best, runner-up, id per query and slot) go to global memory; k_wrap_commit merges.
template <int D>
static __global__ void __launch_bounds__(128) k_wrap_scan(Dev<D> dv, WrapDev<D> wd, int cur, int tb) {
    constexpr int S = X32<D>::STRIDE, TP = HVB_WRAP_TP, NS = HVB_WRAP_NS;
    __shared__ __align__(128) float tile[NS][TP * S];
    __shared__ __align__(8) unsigned long long full[NS];
    __shared__ double m_c1[4][32], m_c2[4][32];
    __shared__ int m_id[4][32];
    const u32 nq = min(__ldcg(wd.nq + cur), wd.wq_cap);
    if (nq == 0) return;
    const WrapShape sh = wrap_shape(nq, dv.n, tb, wd.pcap);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qw = warp % sh.QW, part = warp / sh.QW;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) wrap_mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    LocalStats ls = {};
    u32 it = 0;                                       // tiles this block has consumed: stage = it % NS, parity = (it / NS) & 1
    const int nwork = sh.ntile * sh.PCH;
    for (int wk = blockIdx.x; wk < nwork; wk += gridDim.x) {
        const int qt = wk / sh.PCH, ch = wk - qt * sh.PCH;
        const int p0 = ch * sh.chunk;
        const int npts = max(min(p0 + sh.chunk, dv.n) - p0, 0);
        const int ntl = (npts + TP - 1) / TP;
        const u32 q = (u32)(qt * 32 * sh.QW + qw * 32 + lane);
        const bool live = q < nq;
        const WrapQuery<D>& w = wd.wq[live ? q : 0];
        WrapLane<D> s;
        wrap_lane_init<D>(w, s);
        if (threadIdx.x == 0) {
            for (int t = 0; t < NS && t < ntl; ++t) {
                const int c4 = (min(TP, npts - t * TP) + 3) & ~3;
                const u32 bytes = (u32)(c4 * S * sizeof(float));
                const int st = (int)((it + t) % NS);
                wrap_mbar_expect(&full[st], bytes);
                wrap_bulk_load(tile[st], dv.x32 + (size_t)(p0 + t * TP) * S, bytes, &full[st]);
            }
        }
        for (int t = 0; t < ntl; ++t) {
            const int st = (int)((it + t) % NS);
            wrap_mbar_wait(&full[st], ((it + t) / NS) & 1u);
            const int cnt_t = min(TP, npts - t * TP);
            const int per = ((cnt_t + sh.PW * 4 - 1) / (sh.PW * 4)) * 4;
            const int a = part * per, b = min(a + per, cnt_t);
            if (live && b > a) wrap_scan<D>(dv, w, tile[st] + a * S, p0 + t * TP + a, b - a, wd.tinyA, wd.E32, s, ls);
            __syncthreads();                          // everyone is done with this tile
            if (threadIdx.x == 0 && t + NS < ntl) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const int c4 = (min(TP, npts - (t + NS) * TP) + 3) & ~3;
                const u32 bytes = (u32)(c4 * S * sizeof(float));
                wrap_mbar_expect(&full[st], bytes);
                wrap_bulk_load(tile[st], dv.x32 + (size_t)(p0 + (t + NS) * TP) * S, bytes, &full[st]);
            }
        }
        it += (u32)ntl;
        if (live) wrap_settle<D>(dv, w, wd.tinyA, wd.E32, s, ls);
        // the warps that shared the tiles of these queries merge in shared memory: one partial result per (chunk, query)
        if (part > 0) { m_c1[warp][lane] = s.c1; m_c2[warp][lane] = s.c2; m_id[warp][lane] = s.id1; }
        __syncthreads();
        if (part == 0 && live) {
            for (int pp = 1; pp < sh.PW; ++pp) {
                const int ow = pp * sh.QW + qw;
                wrap_merge(s.c1, s.id1, s.c2, m_c1[ow][lane], m_id[ow][lane], m_c2[ow][lane]);
            }
            const size_t idx = (size_t)ch * nq + q;
            wd.pc1[idx] = s.c1; wd.pc2[idx] = s.c2; wd.pid[idx] = s.id1;
        }
        __syncthreads();
    }
    __syncwarp();
    flush_stats(ls, dv.ctr);
}

// a warp per query: merge of the partial results, then lane 0 builds the new facet
template <int D>
static __global__ void __launch_bounds__(128) k_wrap_commit(Dev<D> dv, HullDev<D> hd, WrapDev<D> wd, int cur, int tb) {
    LocalStats ls = {};
    const u32 nq = min(__ldcg(wd.nq + cur), wd.wq_cap);
    const WrapShape sh = wrap_shape(nq, dv.n, tb, wd.pcap);
    const int slots = sh.PCH;
    const int lane = threadIdx.x & 31;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nq; q += nwarps) {
        double c1 = -INFINITY, c2 = -INFINITY;
        int g = -1;
        for (int sl = lane; sl < slots; sl += 128) {
            // four independent loads in flight per lane (the loop is a chain of memory round trips otherwise)
            double a1[4], a2[4]; int ai[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const bool in = sl + 32 * r < slots;
                const size_t idx = (size_t)(in ? sl + 32 * r : sl) * nq + q;
                a1[r] = __ldcg(wd.pc1 + idx); a2[r] = __ldcg(wd.pc2 + idx); ai[r] = in ? __ldcg(wd.pid + idx) : -1;
                if (!in) { a1[r] = -INFINITY; a2[r] = -INFINITY; }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) wrap_merge(c1, g, c2, a1[r], ai[r], a2[r]);
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const double oc1 = __shfl_xor_sync(0xffffffffu, c1, m), oc2 = __shfl_xor_sync(0xffffffffu, c2, m);
            const int og = __shfl_xor_sync(0xffffffffu, g, m);
            wrap_merge(c1, g, c2, oc1, og, oc2);
        }
        if (lane == 0) wrap_commit<D>(dv, hd, wd, wd.wq[q], c1, g, c2, 1 - cur, ls);
        __syncwarp();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && nq > 0) {
        atomicAdd((u64*)&dv.ctr->raycasts, (u64)nq);
        atomicAdd((u64*)&dv.ctr->cand32, (u64)nq * (u64)dv.n);
        atomicAdd((u64*)&dv.ctr->stages, 1ULL);
    }
    flush_stats(ls, dv.ctr);
}
// facets in caller numbering: edge = the d generators (sorted, 1-based), base = circumcentre of the generators in the
// facet's hyperplane, dir = outward unit normal -- both recomputed from the generators in ascending caller order, so
// that the output does not depend on the path that found the facet
template <int D>
static __global__ void k_final_facets_wrap(Dev<D> dv, HullDev<D> hd, const int* __restrict__ perm, u32 nf,
                               long long* __restrict__ edge, double* __restrict__ base, double* __restrict__ dir, long long* __restrict__ node,
                               u32* __restrict__ out_count, u32* __restrict__ bad) {
    u32 f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int* fs = hd.fsig + (size_t)f * D;
    if (fs[0] < 0) return;
    const u32 o = atomicAdd(out_count, 1u);
    long long e[D];
    int in[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { in[k] = fs[k]; e[k] = (long long)perm[in[k]] + 1; }
    for (int a = 1; a < D; ++a) {
        const long long key = e[a]; const int ki = in[a];
        int b = a - 1;
        while (b >= 0 && e[b] > key) { e[b + 1] = e[b]; in[b + 1] = in[b]; --b; }
        e[b + 1] = key; in[b + 1] = ki;
    }
    double P[D][D], nrm[D], cen[D];
    for (int i = 0; i < D; ++i)
        for (int k = 0; k < D; ++k) P[i][k] = dv.x64[(size_t)in[i] * D + k];
    if (!wrap_facet_geometry<D>(P, hd.fu + (size_t)f * D, nrm, cen)) atomicAdd(bad, 1u);
    for (int k = 0; k < D; ++k) {
        edge[(size_t)o * D + k] = e[k];
        base[(size_t)o * D + k] = cen[k];
        dir[(size_t)o * D + k] = nrm[k];
    }
    node[o] = e[0];
}

// ------------------------------------------------------------------------------------------------------------
// finalize: caller numbering, canonical coordinates, packed sort keys
// ------------------------------------------------------------------------------------------------------------
// one thread per stored record; dead records (lost insertion races) are skipped
template <int D>
static __global__ void k_final_rows(Dev<D> dv, const int* __restrict__ perm, u32 nrec, int bits,
                             long long* __restrict__ out_sig, double* __restrict__ out_r,
                             u64* __restrict__ key_top, u64* __restrict__ key_hi, u64* __restrict__ key_lo, u32* __restrict__ out_count,
                             double* __restrict__ max_var, const unsigned char* __restrict__ owner, int rank, u32 skip_below,
                             double variance_tol, double break_tol, u32* __restrict__ tol_counts, int n_user,
                             double flat_tol, u32* __restrict__ flat_count, int* __restrict__ bucket_count) {
    u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nrec || v < skip_below) return;          // records below skip_below are the caller's own (seed) vertices
    const int* s = dv.vsig + (size_t)v * (D + 1);
    if (s[0] < 0) return;
    // multi-GPU ownership rule: a vertex belongs to the rank that explores the cell of its first generator in grid order
    // (s is sorted by internal = grid-order id); every rank finds all vertices that touch its cells, so the owned sets
    // are disjoint and their union is the full set -- a deterministic dedup that needs no communication
    // (periodic contexts: the first generator that is one of the CALLER's -- halo copies belong to no slab)
    if (owner) {
        int first = -1;
#pragma unroll
        for (int k = D; k >= 0; --k) { const int id = s[k]; if (id < dv.n && perm[id] < n_user) first = id; }
        if (first < 0 || owner[first] != rank) return;
    }
    int in[D + 1];
    long long og[D + 1];
#pragma unroll
    for (int k = 0; k < D + 1; ++k) { in[k] = s[k]; og[k] = (in[k] < dv.n) ? (long long)perm[in[k]] : (long long)in[k]; }
    // sort by caller id (insertion sort network, D + 1 <= 7)
#pragma unroll
    for (int i = 1; i < D + 1; ++i) {
#pragma unroll
        for (int j = i; j > 0; --j) {
            if (og[j - 1] > og[j]) {
                long long t = og[j]; og[j] = og[j - 1]; og[j - 1] = t;
                int ti = in[j]; in[j] = in[j - 1]; in[j - 1] = ti;
            }
        }
    }
    double r[D];
    double flat = 1.0;
    double var = canonical_vertex<D>(dv, in, r, &flat);
    // Non-general position resolved by perturbation (Ctx::resolve_degenerate): the search ran on perturbed generators, the
    // coordinates are solved from the caller's.  A simplex whose d + 1 generators lie in one hyperplane there is a sliver of
    // the perturbed triangulation between two cospherical cells, not a vertex of the caller's diagram: dropped
    if (flat_tol > 0 && !(flat > flat_tol)) { atomicAdd(flat_count, 1u); return; }
    // walkray_correct_vertex (raycast.jl:257-279): the relative variance of the squared radii AFTER the correction decides;
    // above break_tol the vertex is irreparable and dropped (SRI_vertex_irreparable), above variance_tol it is kept
    // and counted (SRI_vertex_suboptimal_correction).  NaN (a singular system) counts as irreparable.
    if (!(var <= break_tol)) { atomicAdd(tol_counts + 0, 1u); return; }
    if (var > variance_tol) atomicAdd(tol_counts + 1, 1u);
    u32 pos = atomicAdd(out_count, 1u);
    if (bucket_count) atomicAdd(bucket_count + (int)og[0], 1);     // rows per first (smallest) generator: sort_rows_bucket
    u64 top = 0, hi = 0, lo = 0;
#pragma unroll
    for (int k = 0; k < D + 1; ++k) {
        out_sig[(size_t)pos * (D + 1) + k] = og[k] + 1;
        // 192-bit key = concatenation of the ids, first id most significant
        top = (top << bits) | (hi >> (64 - bits));
        hi = (hi << bits) | (lo >> (64 - bits));
        lo = (lo << bits) | (u64)og[k];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) out_r[(size_t)pos * D + k] = r[k];
    key_top[pos] = top; key_hi[pos] = hi; key_lo[pos] = lo;
    if (var > 1e-18) atomicMax(reinterpret_cast<unsigned long long*>(max_var), (unsigned long long)__double_as_longlong(var));
}

// every live vertex record of the walk as a row of sorted 1-based caller ids (no filter, no order among rows): the
// complete simplicial vertex set of a perturbed cloud, for the cell volumes of a mesh resolved from non-general position
template <int D>
static __global__ void k_rows_from_records(Dev<D> dv, const int* __restrict__ perm, u32 nrec, long long* __restrict__ out, u32* __restrict__ count) {
    u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nrec) return;
    const int* s = dv.vsig + (size_t)v * (D + 1);
    if (s[0] < 0) return;
    long long og[D + 1];
#pragma unroll
    for (int k = 0; k < D + 1; ++k) og[k] = ((s[k] < dv.n) ? (long long)perm[s[k]] : (long long)s[k]) + 1;
    for (int a = 1; a < D + 1; ++a) { const long long key = og[a]; int b = a - 1; while (b >= 0 && og[b] > key) { og[b + 1] = og[b]; --b; } og[b + 1] = key; }
    const u32 pos = atomicAdd(count, 1u);
    for (int k = 0; k < D + 1; ++k) out[(size_t)pos * (D + 1) + k] = og[k];
}
static __global__ void k_iota(u32* a, u32 n) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}
static __global__ void k_gather_u64(const u64* __restrict__ src, const u32* __restrict__ idx, u64* __restrict__ dst, u32 n) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
// Lexicographic order of the rows for d <= 3 (Ctx::sort_rows_bucket): a counting sort on the first id (always a generator, the
// smallest id of the row) and an insertion sort inside each bucket (6.6 rows on average at d = 3) replace the nine radix passes
// over 68-bit keys.  The row count is a device word: nothing here waits for the host.
static __global__ void k_bucket_scatter(const long long* __restrict__ sig, int stride, const u32* __restrict__ count_ptr, const int* __restrict__ start,
                                        int* __restrict__ cursor, u32* __restrict__ idx) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count_ptr) return;
    const int b = (int)(sig[(size_t)i * stride] - 1);
    idx[start[b] + atomicAdd(cursor + b, 1)] = i;
}
// inside a bucket the rows share their first id, so the low 64 bits of the packed sort key (the remaining d ids, d * bits <= 64)
// order them: keys and row numbers of a bucket are sorted in thread-local arrays (buckets of up to 48 rows; larger ones in place)
static __global__ void k_bucket_sort(const u64* __restrict__ key, const int* __restrict__ start, int nb, u32* __restrict__ idx) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int lo = start[b], m = start[b + 1] - lo;
    if (m <= 1) return;
    if (m <= 48) {
        u64 k[48];
        u32 v[48];
        for (int a = 0; a < m; ++a) { v[a] = idx[lo + a]; k[a] = key[v[a]]; }
        for (int a = 1; a < m; ++a) {
            const u64 kk = k[a]; const u32 vv = v[a];
            int j = a - 1;
            while (j >= 0 && k[j] > kk) { k[j + 1] = k[j]; v[j + 1] = v[j]; --j; }
            k[j + 1] = kk; v[j + 1] = vv;
        }
        for (int a = 0; a < m; ++a) idx[lo + a] = v[a];
        return;
    }
    for (int a = 1; a < m; ++a) {
        const u32 vv = idx[lo + a];
        const u64 kk = key[vv];
        int j = a - 1;
        while (j >= 0 && key[idx[lo + j]] > kk) { idx[lo + j + 1] = idx[lo + j]; --j; }
        idx[lo + j + 1] = vv;
    }
}
template <int D>
static __global__ void k_gather_rows(const long long* __restrict__ sig_in, const double* __restrict__ r_in, const u32* __restrict__ idx,
                              long long* __restrict__ sig_out, double* __restrict__ r_out, u32 n, const u32* __restrict__ n_ptr = nullptr) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_ptr) n = *n_ptr;
    if (i >= n) return;
    u32 s = idx[i];
#pragma unroll
    for (int k = 0; k < D + 1; ++k) sig_out[(size_t)i * (D + 1) + k] = sig_in[(size_t)s * (D + 1) + k];
#pragma unroll
    for (int k = 0; k < D; ++k) r_out[(size_t)i * D + k] = r_in[(size_t)s * D + k];
}

// unbounded edges in caller numbering (pushray!, abstractmesh.jl:191; node = exploring cell = smallest id of the edge)
template <int D>
static __global__ void k_final_rays(Dev<D> dv, const int* __restrict__ perm, u32 nrays,
                             long long* __restrict__ edge, double* __restrict__ base, double* __restrict__ dir, long long* __restrict__ node,
                             const unsigned char* __restrict__ owner, int rank, u32* __restrict__ out_count) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrays) return;
    u32 it = dv.ray_item[i];
    u32 v = it >> 3; int kd = it & 7;
    long long e[D];
    int c = 0, first = 0x7fffffff;
    for (int k = 0; k < D + 1; ++k) {
        if (k == kd) continue;
        int id = dv.vsig[(size_t)v * (D + 1) + k];
        first = min(first, id);
        e[c++] = ((id < dv.n) ? (long long)perm[id] : (long long)id) + 1;
    }
    // multi-GPU: an unbounded edge belongs to the rank of its first generator in grid order (every rank whose cells
    // touch the edge walks it)
    u32 o = i;
    if (owner) {
        if (owner[first] != rank) return;
        o = atomicAdd(out_count, 1u);
    }
    for (int a = 1; a < D; ++a) { long long key = e[a]; int b = a - 1; while (b >= 0 && e[b] > key) { e[b + 1] = e[b]; --b; } e[b + 1] = key; }
    for (int k = 0; k < D; ++k) {
        edge[(size_t)o * D + k] = e[k];
        base[(size_t)o * D + k] = dv.vr[(size_t)v * D + k];
        dir[(size_t)o * D + k] = dv.ray_u[(size_t)i * D + k];
    }
    node[o] = e[0];
}

// ------------------------------------------------------------------------------------------------------------
// wire format of the multi-GPU exchange and of the compact host interface: a row is (D + 1) int32 ids + D doubles
// ------------------------------------------------------------------------------------------------------------
static __global__ void k_narrow_i64(const long long* __restrict__ src, int* __restrict__ dst, size_t count) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < count) dst[i] = (int)src[i];
}
// gathered segments (nseg x cap rows, seg_count[k] valid rows each) -> one dense int64 / double row array
template <int D>
static __global__ void k_unpack_segments(const int* __restrict__ sig32, const double* __restrict__ r_in, int nseg, u32 cap,
                                  const long long* __restrict__ seg_count, long long* __restrict__ sig_out, double* __restrict__ r_out) {
    const u32 row = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (row >= cap || row >= (u32)seg_count[k]) return;
    long long at = 0;
    for (int j = 0; j < k; ++j) at += seg_count[j];
    const size_t src = (size_t)k * cap + row, dst = (size_t)at + row;
#pragma unroll
    for (int c = 0; c < D + 1; ++c) sig_out[dst * (D + 1) + c] = (long long)sig32[src * (D + 1) + c];
#pragma unroll
    for (int c = 0; c < D; ++c) r_out[dst * D + c] = r_in[src * D + c];
}

// ------------------------------------------------------------------------------------------------------------
// neighbour lists (neighbors_of_cell_new, neighbors.jl:219-262): set of unordered id pairs -> CSR
// ------------------------------------------------------------------------------------------------------------
// inserts the unordered pairs of one vertex (s ascending, 1-based caller ids, planes last) into the pair set.
// `own` (may be null = every cell): byte mask over caller cells; only lists of owned cells are built (multi-GPU slabs:
// a rank finds every vertex of its own cells, so these lists are complete; the lists of other cells would be partial
// and are left empty)
template <int D>
__device__ __forceinline__ void pairs_of_row(const long long (&s)[D + 1], long long n, u64* __restrict__ ptab, u64 pmask,
                                             u32* __restrict__ deg, u32* __restrict__ flags, const unsigned char* __restrict__ own) {
    bool ow[D + 1];
#pragma unroll
    for (int i = 0; i < D + 1; ++i) ow[i] = (s[i] <= n) && (!own || own[s[i] - 1]);
#pragma unroll
    for (int i = 0; i < D + 1; ++i) {
#pragma unroll
        for (int j = i + 1; j < D + 1; ++j) {
            if (!(ow[i] || ow[j])) continue;              // neither is a cell whose list is built (ids ascend: planes / halo last)
            u64 key = ((u64)s[i] << 32) | (u64)s[j];      // 1-based ids: key != 0
            u64 slot = mix64(key) & pmask;
            for (u32 probe = 0;; ++probe) {
                u64 cur = __ldcg(ptab + slot);
                if (cur == key) break;
                if (cur == 0) {
                    cur = atomicCAS(ptab + slot, 0ULL, key);
                    if (cur == 0) {
                        if (ow[i]) atomicAdd(deg + (s[i] - 1), 1u);
                        if (ow[j]) atomicAdd(deg + (s[j] - 1), 1u);
                        break;
                    }
                    if (cur == key) break;
                }
                slot = (slot + 1) & pmask;
                if (probe > pmask) { atomicOr(flags, 8u); break; }
            }
        }
    }
}

template <int D>
static __global__ void k_pairs(const long long* __restrict__ sig, u32 nv, long long n, u64* __restrict__ ptab, u64 pmask,
                        u32* __restrict__ deg, u32* __restrict__ flags, const unsigned char* __restrict__ own) {
    u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    long long s[D + 1];
#pragma unroll
    for (int k = 0; k < D + 1; ++k) s[k] = sig[(size_t)v * (D + 1) + k];
    pairs_of_row<D>(s, n, ptab, pmask, deg, flags, own);
}

// the same straight from the vertex records of the walk (internal ids; dead records skipped): the lists do not have to
// wait for the result rows, so they are built next to k_final_rows and the row sort
template <int D>
static __global__ void k_pairs_raw(Dev<D> dv, const int* __restrict__ perm, u32 nrec, long long n, u64* __restrict__ ptab, u64 pmask,
                            u32* __restrict__ deg, u32* __restrict__ flags, const unsigned char* __restrict__ own) {
    u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nrec) return;
    const int* sp = dv.vsig + (size_t)v * (D + 1);
    if (sp[0] < 0) return;
    long long s[D + 1];
#pragma unroll
    for (int k = 0; k < D + 1; ++k) { const int id = sp[k]; s[k] = ((id < dv.n) ? (long long)perm[id] : (long long)id) + 1; }
#pragma unroll
    for (int i = 1; i < D + 1; ++i) {
#pragma unroll
        for (int j = i; j > 0; --j)
            if (s[j - 1] > s[j]) { const long long t = s[j]; s[j] = s[j - 1]; s[j - 1] = t; }
    }
    pairs_of_row<D>(s, n, ptab, pmask, deg, flags, own);
}

static __global__ void k_pair_fill(const u64* __restrict__ ptab, u64 nslots, long long n, const long long* __restrict__ off,
                            u32* __restrict__ cursor, long long* __restrict__ ids, const unsigned char* __restrict__ own) {
    u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    u64 key = ptab[i];
    if (key == 0) return;
    long long a = (long long)(key >> 32), b = (long long)(key & 0xffffffffULL);
    if (a <= n && (!own || own[a - 1])) ids[off[a - 1] + atomicAdd(cursor + (a - 1), 1u)] = b;
    if (b <= n && (!own || own[b - 1])) ids[off[b - 1] + atomicAdd(cursor + (b - 1), 1u)] = a;
}

// own[c] = 1 for the caller cells this context explores
static __global__ void k_own_mask(const int* __restrict__ perm, int n, const unsigned char* __restrict__ owner, int rank, unsigned char* __restrict__ own) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) own[perm[i]] = (!owner || owner[i] == rank) ? 1 : 0;
}

static __global__ void k_sort_lists(const long long* __restrict__ off, long long* __restrict__ ids, long long n) {
    long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= n) return;
    long long a = off[c], b = off[c + 1];
    for (long long i = a + 1; i < b; ++i) {
        long long key = ids[i];
        long long j = i - 1;
        while (j >= a && ids[j] > key) { ids[j + 1] = ids[j]; --j; }
        ids[j + 1] = key;
    }
}

static __global__ void k_u32_to_i64(const u32* __restrict__ a, long long* __restrict__ b, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) b[i] = a[i];
}

// ------------------------------------------------------------------------------------------------------------
// refinement: clean_affected! (meshrefine.jl:126-149) on the device.  The context holds ALL generators
// (old and new); a vertex of the old mesh survives iff no NEW generator lies inside its ball: the reference keeps
// it iff |x_sig1 - r| <= (1 + 1e-7) * dist(r, nearest new node).  Generators of removed vertices are marked affected.
// ------------------------------------------------------------------------------------------------------------
template <int D>
static __global__ void k_clean_affected(Dev<D> dv, const int* __restrict__ perm, const long long* __restrict__ sig, const double* __restrict__ r,
                                 long long nv, int stride, const double* __restrict__ xs, long long new_lo, long long new_hi,
                                 unsigned char* __restrict__ keep, unsigned char* __restrict__ affected, u32* __restrict__ bad) {
    long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const long long* s = sig + v * stride;
    long long g0 = 0;
    for (int k = 0; k < stride && g0 == 0; ++k) if (s[k] >= 1 && s[k] <= dv.n) g0 = s[k];
    if (g0 == 0) { atomicAdd(bad, 1u); keep[v] = 0; return; }        // a vertex names at least one generator
    RayQ<D> q;
    double cen[D], R2 = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) { cen[k] = r[v * D + k]; const double t = xs[(size_t)(g0 - 1) * D + k] - cen[k]; R2 += t * t; q.u[k] = 0.0; q.x0[k] = cen[k]; }
    const double rho = sqrt(R2);
    int clo[D], chi[D], nrows = 1;
    bool empty = false;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const double gk = (double)dv.g[k];
        const double vlo = fmin(fmax((cen[k] - rho - dv.lo[k]) * dv.inv_h[k] - 1e-9, 0.0), gk - 1.0);
        const double vhi = fmin(fmax((cen[k] + rho - dv.lo[k]) * dv.inv_h[k] + 1e-9, -1.0), gk - 1.0);
        clo[k] = (int)floor(vlo); chi[k] = (int)floor(vhi);
        if (chi[k] < clo[k]) empty = true;
        if (k < D - 1) nrows *= (chi[k] - clo[k] + 1);
    }
    bool invaded = false;
    const double lim = R2 / ((1.0 + 1e-7) * (1.0 + 1e-7));
    for (int j = 0; j < nrows && !empty && !invaded; ++j) {
        int pa = 0, pb = 0;
        if (!row_range<D>(dv, q, clo, chi, cen, R2 * (1.0 + 1e-12) + 1e-300, j, pa, pb)) continue;     // u = 0: no half-space clipping
        for (int p = pa; p < pb && !invaded; ++p) {
            const long long cid = perm[p];
            if (cid < new_lo || cid >= new_hi) continue;
            double d2 = 0;
#pragma unroll
            for (int k = 0; k < D; ++k) { const double t = dv.x64[(size_t)p * D + k] - cen[k]; d2 += t * t; }
            invaded = d2 < lim;
        }
    }
    keep[v] = invaded ? 0 : 1;
    if (invaded)
        for (int k = 0; k < stride; ++k) if (s[k] >= 1 && s[k] <= dv.n) affected[s[k] - 1] = 1;
}

// ------------------------------------------------------------------------------------------------------------
// cell volumes from the result rows (hvb_geometry.cuh).  One thread per row; the d+1 contributions of a row are added
// in 64-bit fixed point, so the sums do not depend on the order of the atomics (bitwise reproducible volumes).
// ------------------------------------------------------------------------------------------------------------
template <int D>
static __global__ void k_cell_volumes(const long long* __restrict__ sig, u32 nv, const double* __restrict__ xs, long long n, long long n_list,
                               const PlaneSet* __restrict__ ps, double scale, long long* __restrict__ acc, unsigned char* __restrict__ sat) {
    u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    long long s[D + 1];
#pragma unroll
    for (int k = 0; k < D + 1; ++k) s[k] = sig[(size_t)v * (D + 1) + k];
    for (int k = 0; k < D + 1; ++k) {
        if (s[k] > n_list) continue;                       // planes and halo generators have no cell of their own here
        const double t = vertex_flag_sum<D>(xs, n, ps, s, k) * scale;
        // a single term beyond 2^62 units would saturate the conversion silently (a foot point far outside the cloud:
        // nearly parallel facets, a bounded cell of an unbounded domain): the cell is marked and gets NaN
        if (!(fabs(t) < 4.6e18)) sat[s[k] - 1] = 1;
        atomicAdd(reinterpret_cast<unsigned long long*>(acc + (s[k] - 1)), (unsigned long long)__double2ll_rn(t));
    }
}
static __global__ void k_volumes_finish(const long long* __restrict__ acc, double inv_scale, double* __restrict__ vol, long long n,
                                 const unsigned char* __restrict__ sat) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) vol[i] = sat[i] ? __longlong_as_double(0x7ff8000000000000LL) : (double)acc[i] * inv_scale;
}
// interface areas aligned with the CSR neighbour lists: entry k of cell i's list (ids ascending) receives the (d-1)-volume
// of the facet between the cells of i and ids[k] (a generator, a halo generator or a boundary plane)
__device__ __forceinline__ long long csr_find(const long long* __restrict__ off, const long long* __restrict__ ids, long long cell, long long id) {
    long long lo = off[cell - 1], hi = off[cell];
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (ids[mid] < id) lo = mid + 1; else hi = mid; }
    return (lo < off[cell] && ids[lo] == id) ? lo : -1;
}
template <int D>
static __global__ void k_cell_areas(const long long* __restrict__ sig, u32 nv, const double* __restrict__ xs, long long n, long long n_list,
                             const PlaneSet* __restrict__ ps, const long long* __restrict__ off, const long long* __restrict__ ids,
                             double scale, long long* __restrict__ acc, unsigned char* __restrict__ sat) {
    u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    long long s[D + 1];
#pragma unroll
    for (int k = 0; k < D + 1; ++k) s[k] = sig[(size_t)v * (D + 1) + k];
    for (int k = 0; k < D + 1; ++k) {
        if (s[k] > n_list) continue;
        for (int q = 0; q < D + 1; ++q) {
            if (q == k) continue;
            const long long pos = csr_find(off, ids, s[k], s[q]);
            if (pos < 0) continue;
            const double t = vertex_flag_sum<D>(xs, n, ps, s, k, q) * scale;
            if (!(fabs(t) < 4.6e18)) sat[pos] = 1;
            atomicAdd(reinterpret_cast<unsigned long long*>(acc + pos), (unsigned long long)__double2ll_rn(t));
        }
    }
}
// facets that contain an unbounded edge have no finite area
static __global__ void k_areas_unbounded(const long long* __restrict__ ray_edge, long long nrays, int D, long long n_list,
                                  const long long* __restrict__ off, const long long* __restrict__ ids, double* __restrict__ area) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nrays) return;
    const long long* e = ray_edge + i * D;
    for (int a = 0; a < D; ++a) {
        if (e[a] < 1 || e[a] > n_list) continue;
        for (int b = 0; b < D; ++b) {
            if (a == b) continue;
            const long long pos = csr_find(off, ids, e[a], e[b]);
            if (pos >= 0) area[pos] = INFINITY;
        }
    }
}
// integrals of 1, x_a, x_a x_b over the cells (vertex_flag_moments, hvb_geometry.cuh): acc[cell][NM] in fixed point (one scale per
// degree); sat[cell] = a term left the fixed-point range
template <int D>
static __global__ void k_cell_moments(const long long* __restrict__ sig, u32 nv, const double* __restrict__ xs, long long n, long long n_list,
                               const PlaneSet* __restrict__ ps, double s0, double s1, double s2, long long* __restrict__ acc,
                               unsigned char* __restrict__ sat) {
    const int NM = 1 + D + D * (D + 1) / 2;
    u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    long long s[D + 1];
#pragma unroll
    for (int k = 0; k < D + 1; ++k) s[k] = sig[(size_t)v * (D + 1) + k];
    for (int k = 0; k < D + 1; ++k) {
        if (s[k] > n_list) continue;
        double m[NM];
        vertex_flag_moments<D>(xs, n, ps, s, k, m);
        for (int a = 0; a < NM; ++a) {
            const double t = m[a] * (a == 0 ? s0 : (a < 1 + D ? s1 : s2));
            if (!(fabs(t) < 4.6e18)) sat[s[k] - 1] = 1;
            atomicAdd(reinterpret_cast<unsigned long long*>(acc + (s[k] - 1) * NM + a), (unsigned long long)__double2ll_rn(t));
        }
    }
}
// local (y = x - x_i) -> global coordinates:  int x_a = x_i,a V + M_a,  int x_a x_b = x_i,a x_i,b V + x_i,a M_b + x_i,b M_a + M_ab
template <int D>
static __global__ void k_moments_finish(const long long* __restrict__ acc, const double* __restrict__ xs, long long n_list,
                                 double i0, double i1, double i2, double* __restrict__ vol, double* __restrict__ first, double* __restrict__ second,
                                 const unsigned char* __restrict__ sat) {
    const int NM = 1 + D + D * (D + 1) / 2;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_list) return;
    if (sat[i]) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        if (vol) vol[i] = nan;
        if (first) for (int k = 0; k < D; ++k) first[(size_t)i * D + k] = nan;
        if (second) for (int k = 0; k < D * D; ++k) second[(size_t)i * D * D + k] = nan;
        return;
    }
    const long long* a = acc + i * NM;
    const double V = (double)a[0] * i0;
    double x[D], m1[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { x[k] = xs[(size_t)i * D + k]; m1[k] = (double)a[1 + k] * i1; }
    if (vol) vol[i] = V;
    if (first) for (int k = 0; k < D; ++k) first[(size_t)i * D + k] = x[k] * V + m1[k];
    if (second) {
        int q = 0;
        for (int k = 0; k < D; ++k)
            for (int l = k; l < D; ++l, ++q) {
                const double val = x[k] * x[l] * V + x[k] * m1[l] + x[l] * m1[k] + (double)a[1 + D + q] * i2;
                second[(size_t)i * D * D + k * D + l] = val;
                second[(size_t)i * D * D + l * D + k] = val;
            }
    }
}
// area and first moment of every interface, aligned with the CSR neighbour lists (vertex_flag_moments with `first`):
// acc[entry][1 + D] in fixed point, sat[entry] = a term left the range
template <int D>
static __global__ void k_cell_area_moments(const long long* __restrict__ sig, u32 nv, const double* __restrict__ xs, long long n, long long n_list,
                                    const PlaneSet* __restrict__ ps, const long long* __restrict__ off, const long long* __restrict__ ids,
                                    double s0, double s1, long long* __restrict__ acc, unsigned char* __restrict__ sat) {
    const int NM = 1 + D + D * (D + 1) / 2;
    u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    long long s[D + 1];
#pragma unroll
    for (int k = 0; k < D + 1; ++k) s[k] = sig[(size_t)v * (D + 1) + k];
    for (int k = 0; k < D + 1; ++k) {
        if (s[k] > n_list) continue;
        for (int q = 0; q < D + 1; ++q) {
            if (q == k) continue;
            const long long pos = csr_find(off, ids, s[k], s[q]);
            if (pos < 0) continue;
            double m[NM];
            vertex_flag_moments<D>(xs, n, ps, s, k, m, q);
            for (int a = 0; a < 1 + D; ++a) {
                const double t = m[a] * (a == 0 ? s0 : s1);
                if (!(fabs(t) < 4.6e18)) sat[pos] = 1;
                atomicAdd(reinterpret_cast<unsigned long long*>(acc + pos * (1 + D) + a), (unsigned long long)__double2ll_rn(t));
            }
        }
    }
}
// local -> global: int x_a over the interface = x_i,a area + M_a; `cell_of[entry]` is the cell whose list holds the entry
template <int D>
static __global__ void k_area_moments_finish(const long long* __restrict__ acc, const double* __restrict__ xs, const long long* __restrict__ off,
                                      long long n_list, long long tot, double i0, double i1, double* __restrict__ area, double* __restrict__ first,
                                      const unsigned char* __restrict__ sat) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= tot) return;
    // the cell of entry e: the last offset <= e
    long long lo = 0, hi = n_list;
    while (lo < hi) { const long long mid = (lo + hi + 1) >> 1; if (off[mid] <= e) lo = mid; else hi = mid - 1; }
    const long long i = lo;                                   // 0-based cell
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double A = (double)acc[e * (1 + D)] * i0;
    if (area) area[e] = sat[e] ? nan : A;
    if (first) for (int k = 0; k < D; ++k) first[(size_t)e * D + k] = sat[e] ? nan : xs[(size_t)i * D + k] * A + (double)acc[e * (1 + D) + 1 + k] * i1;
}
static __global__ void k_area_moments_unbounded(const long long* __restrict__ ray_edge, long long nrays, int D, long long n_list,
                                         const long long* __restrict__ off, const long long* __restrict__ ids, double* __restrict__ area, double* __restrict__ first) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nrays) return;
    const long long* e = ray_edge + i * D;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int a = 0; a < D; ++a) {
        if (e[a] < 1 || e[a] > n_list) continue;
        for (int b = 0; b < D; ++b) {
            if (a == b) continue;
            const long long pos = csr_find(off, ids, e[a], e[b]);
            if (pos < 0) continue;
            if (area) area[pos] = INFINITY;
            if (first) for (int k = 0; k < D; ++k) first[(size_t)pos * D + k] = nan;
        }
    }
}
// cells with an unbounded edge: volume +inf, moments undefined (NaN)
static __global__ void k_moments_unbounded(const long long* __restrict__ ray_edge, long long nentries, long long n_list, int dim,
                                    double* __restrict__ vol, double* __restrict__ first, double* __restrict__ second) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nentries) return;
    const long long g = ray_edge[i];
    if (g < 1 || g > n_list) return;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (vol) vol[g - 1] = INFINITY;
    if (first) for (int k = 0; k < dim; ++k) first[(size_t)(g - 1) * dim + k] = nan;
    if (second) for (int k = 0; k < dim * dim; ++k) second[(size_t)(g - 1) * dim * dim + k] = nan;
}
// cells with an unbounded edge have no finite volume
static __global__ void k_volumes_unbounded(const long long* __restrict__ ray_edge, long long nentries, long long n_list, double* __restrict__ vol) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nentries) return;
    const long long g = ray_edge[i];
    if (g >= 1 && g <= n_list) vol[g - 1] = INFINITY;
}

// ------------------------------------------------------------------------------------------------------------
// multi-GPU merge: dedup of gathered rows (sorted caller ids, 1-based) by a row hash set
// ------------------------------------------------------------------------------------------------------------
template <int D>
static __global__ void k_merge_rows(const long long* __restrict__ sig_in, const double* __restrict__ r_in, u64 count, int bits,
                             u64* __restrict__ tab, u64 mask, long long* __restrict__ sig_out, double* __restrict__ r_out,
                             u64* __restrict__ key_top, u64* __restrict__ key_hi, u64* __restrict__ key_lo, u32* __restrict__ out_count) {
    u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (i >= count) return;
    long long s[D + 1];
    u64 h = 0x9e3779b97f4a7c15ULL;
#pragma unroll
    for (int k = 0; k < D + 1; ++k) { s[k] = sig_in[i * (D + 1) + k]; h = (h ^ (u64)s[k]) * 0x100000001b3ULL + 0x632be59bd9b4e019ULL; }
    h = mix64(h);
    u64 slot = h & mask;
    for (;;) {
        u64 cur = __ldcg(tab + slot);
        if (cur == 0) {
            cur = atomicCAS(tab + slot, 0ULL, i + 1);
            if (cur == 0) break;                          // this row represents its signature
        }
        const long long* o = sig_in + (cur - 1) * (D + 1);
        bool eq = true;
#pragma unroll
        for (int k = 0; k < D + 1; ++k) eq &= (o[k] == s[k]);
        if (eq) return;                                   // duplicate found by another rank
        slot = (slot + 1) & mask;
    }
    u32 pos = atomicAdd(out_count, 1u);
    u64 top = 0, hi = 0, lo = 0;
#pragma unroll
    for (int k = 0; k < D + 1; ++k) {
        sig_out[(size_t)pos * (D + 1) + k] = s[k];
        top = (top << bits) | (hi >> (64 - bits));
        hi = (hi << bits) | (lo >> (64 - bits));
        lo = (lo << bits) | (u64)(s[k] - 1);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) r_out[(size_t)pos * D + k] = r_in[i * D + k];
    key_top[pos] = top; key_hi[pos] = hi; key_lo[pos] = lo;
}

}  // namespace hvb

#include "hvb_coop.cuh"
