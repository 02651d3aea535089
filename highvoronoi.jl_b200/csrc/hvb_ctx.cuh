#pragma once
// hvb_ctx.cuh -- host orchestration of the B200 raycast vertex search: one context = one GPU (Ctx<D>).
// Compiled once per dimension (hvb_dim.cu, -DHVB_DIM=2..6) so that the five instantiations build in parallel; the C ABI
// (hvb_api.cu) only sees the abstract hvb_ctx.
//
// One context = one GPU.  hvb_create uploads the generators and builds the uniform-grid index; hvb_search seeds
// the frontier (descents), runs frontier rounds until no open edge is left, re-seeds cells that are still empty,
// then finalizes (caller numbering, canonical coordinates, lexicographic order) and stages the result in
// page-locked host memory.  There is no host compute path: if CUDA is unavailable every entry fails.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hvb200.h"
#include "hvb_host.hpp"
#include "hvb_nongeneral.hpp"
#include "hvb_ctx_base.hpp"
#include "hvb_kernels.cuh"
#include "hvb_nccl.hpp"

using namespace hvb;


#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            char buf_[512];                                                                              \
            snprintf(buf_, sizeof(buf_), "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            this->err = buf_;                                                                            \
            return (e_ == cudaErrorMemoryAllocation) ? HVB_ENOMEM : HVB_ECUDA;                           \
        }                                                                                                \
    } while (0)

#define NK(call)                                                                                         \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != ncclSuccess) {                                                                         \
            char buf_[512];                                                                              \
            snprintf(buf_, sizeof(buf_), "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, hvb::Nccl::get().GetErrorString(r_)); \
            this->err = buf_;                                                                            \
            return HVB_ENCCL;                                                                            \
        }                                                                                                \
    } while (0)

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        // grow geometrically: sizes that creep up from step to step must not re-allocate (cudaFree synchronises)
        size_t want = std::max<size_t>(n + n / 4, 1);
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e != cudaSuccess) { want = std::max<size_t>(n, 1); e = cudaMalloc((void**)&p, want * sizeof(T)); }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // grows to n elements and keeps the first `keep` elements (device-to-device copy on `st`)
    cudaError_t grow_keep(size_t n, size_t keep, cudaStream_t st) {
        if (n <= cap) return cudaSuccess;
        T* np_ = nullptr;
        size_t want = std::max<size_t>(n + n / 4, 1);
        cudaError_t e = cudaMalloc((void**)&np_, want * sizeof(T));
        if (e != cudaSuccess) { want = std::max<size_t>(n, 1); e = cudaMalloc((void**)&np_, want * sizeof(T)); }
        if (e != cudaSuccess) return e;
        if (p && keep) e = cudaMemcpyAsync(np_, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (p) cudaFree(p);
        p = np_; cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <class T>
struct HBuf {   // page-locked host memory
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = std::max<size_t>(n + n / 4, 1);
        cudaError_t e = cudaHostAlloc((void**)&p, want * sizeof(T), cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Round { u32 qcount; u32 cursor; };   // frontier length and the work cursor of the round that consumes it
struct Scalars {            // small device-side words, mirrored into pinned host memory after every round
    Round rnd[2];
    u32 vcount;
    u32 ray_count;
    u32 unseeded;
    u32 out_count;
    u32 pflags;
    u32 q_head, q_done, q_abort;      // single-launch walk (k_walk): tickets, processed entries, safety abort
    u32 bbox_done;
    u32 bbox_viol;
    u32 pad2;
    u32 ray_out;                      // unbounded edges this rank owns (k_final_rays with slabs)
    u32 pad3;
    u32 tol_counts[2];                // vertices dropped (variance > break_tol) / kept above variance_tol (k_final_rows)
    double max_var;
    double bbox[12];
};

// Small device words -> page-locked host memory, written by the SMs (the host pointers are device-accessible under
// unified addressing).  A cudaMemcpy would queue on the D2H copy engine BEHIND the result staging copy that runs on
// the staging stream: the neighbour build then waited a whole staging copy for every 100-byte read.
static __global__ void k_publish(const u32* __restrict__ a, u32* ha, int na, const u32* __restrict__ b, u32* hb, int nb,
                          const long long* __restrict__ extra, long long* hextra) {
    for (int i = threadIdx.x; i < na; i += blockDim.x) ha[i] = __ldcg(a + i);
    for (int i = threadIdx.x; i < nb; i += blockDim.x) hb[i] = __ldcg(b + i);
    if (extra && threadIdx.x == 0) *hextra = __ldcg(extra);
    __threadfence_system();
}


template <int D>
struct Ctx : hvb_ctx {
    int G = 1;                       // lanes per frontier entry; 1 measured best for d = 2..5 (prm.tile_size overrides)
    bool debug = false;
    bool persistent = true;          // single-launch walk (k_walk) instead of one launch per frontier round (prm.persistent)
    bool coop = true;                // prm.persistent == 2: the warp-cooperative query (k_walk_coop), lane-per-ray tiles only
    cudaStream_t stream = nullptr, sstream = nullptr;    // compute stream, result-staging (D2H) stream
    cudaStream_t nstream = nullptr;                      // neighbour lists, built next to the row sort (both read the unsorted rows)
    cudaStream_t sstream2 = nullptr;                     // staging of the neighbour lists (whichever of rows / lists is ready first goes first)
    cudaEvent_t ev_stage = nullptr, ev_nb = nullptr;
    struct NbScalars { u32 pflags; u32 pad; };
    DBuf<NbScalars> nbsc;
    HBuf<NbScalars> h_nbsc;
    HBuf<long long> h_nbtotal;
    int sms = 148;
    Dev<D> dv;
    int64_t ncells = 0;
    // index
    DBuf<double> xs_in, x64;
    DBuf<float> x32;
    DBuf<int> perm, inv, cell_of, cell_start, cell_cur, unseeded_list;
    DBuf<PlaneSet> planes;
    DBuf<unsigned char> active, has_vertex;
    DBuf<char> cub_tmp, nb_cub_tmp;
    DBuf<double> bbox_partial;
    // search state
    int64_t vcap = 0;
    DBuf<int> vsig;
    DBuf<double> vr;
    DBuf<u64> vtab, etab;
    DBuf<u64> q[2];
    u32 qcap = 0;
    DBuf<u32> ray_item;
    DBuf<double> ray_u;
    u32 ray_cap = 0;
    DBuf<Counters> ctr;
    DBuf<Scalars> sc;
    HBuf<Scalars> h_sc;
    HBuf<Counters> h_ctr;
    HBuf<long long> h_extra;
    DBuf<long long> cells_dev, seed_sig_dev;
    DBuf<double> seed_r_dev;
    u32 seed_prefix = 0;              // vertex records [0, seed_prefix) are the caller's own vertices (not returned)
    // results
    DBuf<long long> out_sig[2];
    DBuf<double> out_r[2];
    DBuf<u64> key_top, key_hi, key_lo, key_tmp;
    DBuf<u32> idx[2];
    int res = 0;                     // which of out_sig/out_r holds the final rows
    int64_t nvert = 0, nrays = 0;
    DBuf<long long> ray_edge, ray_node;
    DBuf<double> ray_base, ray_dir;
    HBuf<long long> h_sig;
    HBuf<double> h_r;
    bool have_result = false, staged = false;
    // neighbours
    DBuf<u64> ptab;
    DBuf<u32> deg, ncur;
    DBuf<long long> nb_off, nb_ids;
    HBuf<long long> h_nb_off, h_nb_ids;
    bool nb_staged = false;
    int64_t nb_total = -1;
    // multi-GPU slabs: byte mask over caller cells, 1 = this context owns the cell (its sorted position lies in the slab).
    // Neighbour lists are built for owned cells only (complete there); null = every cell
    DBuf<unsigned char> own_mask;
    const unsigned char* own_ptr = nullptr;
    // multi-GPU decomposition (hvb_kernels.cuh, BlockSpec): rank per sorted position, built with the index
    DBuf<unsigned char> owner;
    const unsigned char* owner_ptr = nullptr;          // null: world == 1
    DBuf<unsigned int> marg;
    HBuf<unsigned int> h_marg;
    BlockSpec bspec;
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t ev_up = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr, ev_s0 = nullptr, ev_s1 = nullptr, ev_n0 = nullptr, ev_n1 = nullptr;
    int64_t launches = 0;

    ~Ctx() override {
        cudaSetDevice(prm.device);
        if (stream) cudaStreamSynchronize(stream);
        xs_in.release(); x64.release(); x32.release(); perm.release(); inv.release(); cell_of.release(); cell_start.release();
        nb_cub_tmp.release(); cell_cur.release(); unseeded_list.release(); bbox_partial.release(); planes.release(); active.release(); has_vertex.release(); cub_tmp.release();
        vsig.release(); vr.release(); vtab.release(); etab.release(); q[0].release(); q[1].release(); ray_item.release(); ray_u.release();
        ctr.release(); sc.release(); h_sc.release(); h_ctr.release(); h_extra.release(); cells_dev.release(); seed_sig_dev.release(); seed_r_dev.release();
        out_sig[0].release(); out_sig[1].release(); out_r[0].release(); out_r[1].release(); key_top.release(); key_hi.release(); key_lo.release(); key_tmp.release();
        idx[0].release(); idx[1].release(); ray_edge.release(); ray_node.release(); ray_base.release(); ray_dir.release();
        bkt_count.release(); bkt_start.release(); xs_orig.release(); xcan.release();
        w_wq.release(); w_pc1.release(); w_pc2.release(); w_pid.release(); w_q[0].release(); w_q[1].release(); w_words.release(); w_seed.release(); hw_words.release();
        h_sig.release(); h_r.release(); sig32_dev.release(); ids32_dev.release(); h_sig32.release(); h_ids32.release(); h_fsig.release(); h_fitem.release(); h_fu.release(); h_ftab.release(); h_rtab.release(); h_arg.release(); h_nb_off.release(); h_nb_ids.release(); ptab.release(); deg.release(); ncur.release(); nb_off.release(); nb_ids.release();
        for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
        if (ev_up) cudaEventDestroy(ev_up);
        if (ev_a) cudaEventDestroy(ev_a);
        if (ev_b) cudaEventDestroy(ev_b);
        if (ev_c) cudaEventDestroy(ev_c);
        if (ev_d) cudaEventDestroy(ev_d);
        if (ev_sd) cudaEventDestroy(ev_sd);
        if (ev_sd2) cudaEventDestroy(ev_sd2);
        if (ev_s0) cudaEventDestroy(ev_s0);
        if (ev_s1) cudaEventDestroy(ev_s1);
        if (ev_n0) cudaEventDestroy(ev_n0);
        if (ev_n1) cudaEventDestroy(ev_n1);
        if (ev_stage) cudaEventDestroy(ev_stage);
        if (ev_nb) cudaEventDestroy(ev_nb);
        nbsc.release(); h_nbsc.release(); h_nbtotal.release(); own_mask.release(); owner.release(); marg.release(); h_marg.release();
        if (comm && comm_owned) Nccl::get().CommDestroy(comm);
        xc_counts.release(); h_xc_counts.release(); xc_counts32.release(); h_xc_counts32.release();
        if (xstream) { cudaStreamDestroy(xstream); cudaEventDestroy(ev_x0); cudaEventDestroy(ev_x1); } xs_sig32.release(); xr_sig32.release(); xs_r.release(); xr_r.release(); xc_red.release(); h_xc_red.release();
        if (nstream) { cudaStreamSynchronize(nstream); cudaStreamDestroy(nstream); }
        if (sstream2) { cudaStreamSynchronize(sstream2); cudaStreamDestroy(sstream2); }
        if (ev_p0) cudaEventDestroy(ev_p0);
        if (ev_p1) cudaEventDestroy(ev_p1);
        vol_acc.release(); vol_dev.release(); mom_dev.release(); vol_sat.release(); ca_keep.release(); ca_aff.release();
        halo_cnt.release(); halo_off.release(); halo_origin.release(); halo_mult.release(); vflags.release(); cert.release(); h_cert.release();
        if (sstream) { cudaStreamSynchronize(sstream); cudaStreamDestroy(sstream); }
        if (stream) cudaStreamDestroy(stream);
    }

    static int blocks_for(int64_t items, int per_block) { return (int)std::max<int64_t>(1, (items + per_block - 1) / per_block); }

    PlaneSet ps_host;                 // the planes on the device (periodic planes pushed outwards by `margin`)
    PlaneSet ps_orig;                 // the caller's planes
    bool setup_done = false;
    // periodic domains: caller generators [0, n_user) + halo copies [n_user, n)
    bool periodic = false;
    int64_t n_user = 0, n_halo = 0;
    double margin = 0;
    HaloSpec halo;
    PeriodicCert pcert;
    double pair_width[HVB_MAX_PAIRS];
    DBuf<int> halo_cnt, halo_off, halo_origin;
    DBuf<signed char> halo_mult;
    DBuf<unsigned char> vflags;
    DBuf<CertOut> cert;
    HBuf<CertOut> h_cert;
    bool have_flags = false;
    double margin_spacing = 0, margin_rel_certified = 0;
    cudaEvent_t ev_p0 = nullptr, ev_p1 = nullptr;

    int init(const double* xs, const double* pbase, const double* pnormal, const int32_t* plane_bc) override {
        memset(&dv, 0, sizeof(dv));
        memset(&st, 0, sizeof(st));
        CK(cudaSetDevice(prm.device));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, prm.device));
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&sstream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&sstream2, cudaStreamNonBlocking));
        {   // the neighbour build yields to the compute stream (the row sort gates the large staging copy)
            int prio_lo = 0, prio_hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            CK(cudaStreamCreateWithPriority(&nstream, cudaStreamNonBlocking, prio_lo));
        }
        CK(cudaEventCreateWithFlags(&ev_stage, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_nb, cudaEventDisableTiming));
        CK(nbsc.ensure(1)); CK(h_nbsc.ensure(1)); CK(h_nbtotal.ensure(1));
        CK(cudaEventCreate(&ev_a)); CK(cudaEventCreate(&ev_b)); CK(cudaEventCreate(&ev_c)); CK(cudaEventCreate(&ev_d));
        CK(cudaEventCreate(&ev_up));
        CK(cudaEventCreate(&ev_s0)); CK(cudaEventCreate(&ev_s1)); CK(cudaEventCreate(&ev_n0)); CK(cudaEventCreate(&ev_n1));
        CK(cudaEventCreate(&ev_p0)); CK(cudaEventCreate(&ev_p1));
        // planes: unit outward normals, offsets
        memset(&ps_host, 0, sizeof(ps_host));
        ps_host.P = P;
        for (int p = 0; p < P; ++p) {
            double nr = 0;
            for (int k = 0; k < D; ++k) nr += pnormal[p * D + k] * pnormal[p * D + k];
            nr = sqrt(nr);
            if (!(nr > 0)) { err = "boundary plane with zero normal"; return HVB_EINVAL; }
            double off = 0;
            for (int k = 0; k < D; ++k) { ps_host.normal[p * 6 + k] = pnormal[p * D + k] / nr; off += ps_host.normal[p * 6 + k] * pbase[p * D + k]; }
            ps_host.off[p] = off;
        }
        ps_orig = ps_host;
        // periodic plane pairs (Plane.BC > 0 names the partner plane, boundary.jl:15-29; cuboid boundary.jl:510-534)
        memset(&halo, 0, sizeof(halo));
        memset(&pcert, 0, sizeof(pcert));
        pcert.nplanes = P;
        for (int p = 0; p < P; ++p) pcert.off_orig[p] = ps_orig.off[p];
        periodic = false;
        if (plane_bc) {
            for (int p = 0; p < P; ++p) {
                int q = plane_bc[p] - 1;
                if (plane_bc[p] <= 0) continue;
                if (q >= P || q == p || plane_bc[q] - 1 != p) { err = "periodic planes must name each other as partners"; return HVB_EINVAL; }
                double dotn = 0;
                for (int k = 0; k < D; ++k) dotn += ps_orig.normal[p * 6 + k] * ps_orig.normal[q * 6 + k];
                if (!(dotn < -1.0 + 1e-9)) { err = "periodic partner planes must be parallel with opposite normals"; return HVB_EINVAL; }
                pcert.is_periodic[p] = 1;
                if (q < p) continue;                      // the pair is recorded once, at its lower plane index
                if (halo.npairs == HVB_MAX_PAIRS) { err = "too many periodic plane pairs"; return HVB_EINVAL; }
                const int i = halo.npairs++;
                halo.plane_a[i] = p; halo.plane_b[i] = q;
                const double width = ps_orig.off[p] + ps_orig.off[q];
                if (!(width > 0)) { err = "periodic planes enclose an empty slab"; return HVB_EINVAL; }
                pair_width[i] = width;
                for (int k = 0; k < D; ++k) halo.T[i][k] = ps_orig.normal[p * 6 + k] * width;
            }
            periodic = halo.npairs > 0;
        }
        CK(planes.ensure(1)); CK(ctr.ensure(1)); CK(sc.ensure(1)); CK(h_sc.ensure(1)); CK(h_ctr.ensure(1)); CK(h_extra.ensure(1));
        CK(cert.ensure(1)); CK(h_cert.ensure(1));
        CK(cudaMemcpyAsync(planes.p, &ps_orig, sizeof(ps_orig), cudaMemcpyHostToDevice, stream));
        dv.plane_tol = prm.plane_tolerance; dv.t_min = prm.plane_tolerance;
        dv.probe_scale = prm.probe_scale > 1.0 ? prm.probe_scale : default_probe_scale(D);
        dv.fp32_filter = prm.fp32_filter;
        dv.probe_growth = 2.0;
        if (prm.tile_size == 1 || prm.tile_size == 2 || prm.tile_size == 4 || prm.tile_size == 8 || prm.tile_size == 16 || prm.tile_size == 32) G = prm.tile_size;
        debug = getenv("HVB_DEBUG") != nullptr;
        if (const char* e = getenv("HVB_PERSISTENT")) prm.persistent = atoi(e);      // tuning / test runs: force a walk variant
        persistent = prm.persistent != 0;
        coop = prm.persistent >= 2;
        setup_done = true;
        return set_points(n, xs);
    }

    // bounding box + domain check (check_boundary, boundary.jl:437) of `cnt` points against the planes on the device;
    // the box lands in h_sc.p->bbox.  `host_xs` (may be null) is only used to word the error message.
    int check_points(const double* dev_xs, int64_t cnt, const double* host_xs) {
        const int bb_blocks = std::min(blocks_for(cnt, 256), sms * 8);
        CK(bbox_partial.ensure((size_t)bb_blocks * 2 * D));
        CK(cudaMemsetAsync(&sc.p->bbox_done, 0, sizeof(u32), stream));
        CK(cudaMemsetAsync(&sc.p->bbox_viol, 0xff, sizeof(u32), stream));
        k_bbox_check<D><<<bb_blocks, 256, 0, stream>>>(dev_xs, (int)cnt, planes.p, bbox_partial.p, &sc.p->bbox_done, sc.p->bbox, &sc.p->bbox_viol);
        ++launches;
        CK(cudaMemcpyAsync(h_sc.p, sc.p, sizeof(Scalars), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        if (h_sc.p->bbox_viol != 0xffffffffu) {
            char b[200];
            if (!host_xs) { snprintf(b, sizeof(b), "internal error: halo generator %u outside the pushed planes", h_sc.p->bbox_viol + 1); err = b; return HVB_ECUDA; }
            const double* x = host_xs + (size_t)h_sc.p->bbox_viol * D;
            bool finite = true;
            for (int k = 0; k < D; ++k) finite &= (x[k] == x[k]) && fabs(x[k]) <= 1e150;
            if (!finite) snprintf(b, sizeof(b), "non-finite coordinate in generator %u", h_sc.p->bbox_viol + 1);
            else {
                int pbad = 0;
                for (int p = 0; p < P; ++p) {
                    double sdot = 0;
                    for (int k = 0; k < D; ++k) sdot += ps_orig.normal[p * 6 + k] * x[k];
                    if (sdot > ps_orig.off[p]) { pbad = p + 1; break; }
                }
                snprintf(b, sizeof(b), "generator %u does not lie in the domain (plane %d)", h_sc.p->bbox_viol + 1, pbad);
            }
            err = b;
            return HVB_EINVAL;
        }
        return HVB_OK;
    }

    // (re)loads the generators and rebuilds the spatial index; every buffer is reused when it is large enough
    int set_points(int64_t n_new, const double* xs) override {
        if (!setup_done) { err = "context not initialised"; return HVB_ESTATE; }
        if (n_new <= D || !xs) { err = "There are not enough points to create a Voronoi tessellation"; return HVB_EINVAL; }
        CK(cudaSetDevice(prm.device));
        have_result = false; staged = false; staged_r = staged_sig64 = staged_sig32 = false; nb_staged32 = false; nb_total = -1; have_flags = false;
        { int b = 1; while ((1LL << b) < n_new + P + 1) ++b;
          if ((D + 1) * b > 192) { err = "too many generators for the lexicographic row order of this dimension (ids need (dim+1) x bits <= 192)"; return HVB_EINVAL; } }
        n_user = n_new; n = n_new; n_halo = 0;
        perturbed = false; merged = false; m_nb_built = false;
        dv.t_min = prm.plane_tolerance;
        CK(cudaEventRecord(ev_a, stream));
        // upload, then bounding box + domain check on the device against the caller's planes
        CK(xs_in.ensure((size_t)n * D));
        CK(cudaMemcpyAsync(xs_in.p, xs, (size_t)n * D * sizeof(double), cudaMemcpyHostToDevice, stream));
        CK(cudaEventRecord(ev_up, stream));              // ms_upload: the copy alone (pageable caller memory makes it slow)
        if (periodic) CK(cudaMemcpyAsync(planes.p, &ps_orig, sizeof(ps_orig), cudaMemcpyHostToDevice, stream));
        int rc = check_points(xs_in.p, n, xs); if (rc) return rc;
        if (periodic) {
            // first margin: twice the typical circumradius of a Poisson-Delaunay simplex at this density, times a
            // dimension-dependent allowance for the largest ball; the certificate (certify()) corrects it if needed
            static const double cd[7] = {0, 0, 3.14159265358979, 4.18879020478639, 4.93480220054468, 5.26378901391432, 5.16771278004997};
            static const double allow[7] = {0, 0, 3.2, 2.4, 1.9, 1.7, 1.55};
            double vol = 1.0;
            for (int k = 0; k < D; ++k) vol *= std::max(h_sc.p->bbox[D + k] - h_sc.p->bbox[k], 1e-300);
            const double spacing = pow(vol / (double)n_user, 1.0 / D);
            margin = prm.periodic_margin > 0 ? prm.periodic_margin : 2.0 * spacing * pow((double)D / cd[D], 1.0 / D) * allow[D];
            // a margin the certificate had to enlarge on an earlier cloud of this context is not tried again
            margin_spacing = spacing;
            if (margin_rel_certified > 0) margin = std::max(margin, margin_rel_certified * spacing);
            rc = build_halo(); if (rc) return rc;
        }
        return build_index();
    }

    // periodic domains: pushes the periodic planes outwards by `margin`, appends the halo copies of the caller's
    // generators that fall inside the pushed planes (reflect_nodes, domain.jl:338) and recomputes the bounding box
    int build_halo() {
        double wmin = 1e300;
        for (int i = 0; i < halo.npairs; ++i) wmin = std::min(wmin, pair_width[i]);
        if (!(margin > 0)) margin = 0.25 * wmin;
        if (margin > 2.0 * wmin) { err = "periodic cells extend over more than two periods: too few generators for this domain"; return HVB_EINCOMPLETE; }
        ps_host = ps_orig;
        halo.ncodes = 1;
        for (int i = 0; i < halo.npairs; ++i) {
            ps_host.off[halo.plane_a[i]] += margin; ps_host.off[halo.plane_b[i]] += margin;
            halo.K[i] = std::max(1, (int)ceil(margin / pair_width[i]));
            halo.ncodes *= 2 * halo.K[i] + 1;
        }
        CK(cudaMemcpyAsync(planes.p, &ps_host, sizeof(ps_host), cudaMemcpyHostToDevice, stream));
        CK(halo_cnt.ensure(n_user + 1)); CK(halo_off.ensure(n_user + 1));
        CK(cudaMemsetAsync(halo_cnt.p + n_user, 0, sizeof(int), stream));
        k_halo<D><<<blocks_for(n_user, 128), 128, 0, stream>>>(xs_in.p, (int)n_user, halo, planes.p, halo_cnt.p, nullptr, nullptr, nullptr, nullptr);
        ++launches;
        size_t tmp_bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, halo_cnt.p, halo_off.p, (int)(n_user + 1), stream));
        CK(cub_tmp.ensure(tmp_bytes));
        CK(cub::DeviceScan::ExclusiveSum(cub_tmp.p, tmp_bytes, halo_cnt.p, halo_off.p, (int)(n_user + 1), stream));
        int total = 0;
        CK(cudaMemcpyAsync(&total, halo_off.p + n_user, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        n_halo = total;
        n = n_user + n_halo;
        if (n > 0x7ff00000LL) { err = "too many generators after adding the periodic halo"; return HVB_EINVAL; }
        CK(xs_in.grow_keep((size_t)n * D, (size_t)n_user * D, stream));
        CK(halo_origin.ensure(std::max<int64_t>(n_halo, 1))); CK(halo_mult.ensure(std::max<int64_t>(n_halo, 1) * halo.npairs));
        if (n_halo > 0) {
            k_halo<D><<<blocks_for(n_user, 128), 128, 0, stream>>>(xs_in.p, (int)n_user, halo, planes.p, nullptr, halo_off.p, xs_in.p, halo_origin.p, halo_mult.p);
            ++launches;
        }
        if (debug) fprintf(stderr, "[hvb] periodic: margin %.4g, %lld halo generators for %lld caller generators\n", margin, (long long)n_halo, (long long)n_user);
        return check_points(xs_in.p, n, nullptr);
    }

    // grid index over the n generators in xs_in (bounding box in h_sc.p->bbox)
    int build_index() {
        dv.n = (int)n;
        double blo[D], bhi[D];
        for (int k = 0; k < D; ++k) { blo[k] = h_sc.p->bbox[k]; bhi[k] = h_sc.p->bbox[D + k]; }
        int ppc = prm.points_per_cell > 0 ? prm.points_per_cell : default_points_per_cell(D);
        ncells = setup_grid<D>(dv, blo, bhi, n, ppc);
        CK(x64.ensure((size_t)n * D)); CK(x32.ensure((size_t)(n + 4) * X32<D>::STRIDE));      // + 4: the hull's bulk copies read whole groups of four generators
        CK(perm.ensure(n)); CK(inv.ensure(n)); CK(cell_of.ensure(n)); CK(unseeded_list.ensure(n));
        CK(cell_start.ensure(ncells + 1)); CK(cell_cur.ensure(ncells + 1));
        CK(active.ensure(n)); CK(has_vertex.ensure(n));
        dv.cell_start = cell_start.p; dv.x32 = x32.p; dv.x64 = x64.p; dv.xcan = x64.p; dv.planes = planes.p; dv.active = active.p;
        dv.has_vertex = has_vertex.p; dv.ctr = ctr.p;
        // counting sort into cells
        CK(cudaMemsetAsync(cell_cur.p, 0, (size_t)(ncells + 1) * sizeof(int), stream));
        k_cell_count<D><<<blocks_for(n, 256), 256, 0, stream>>>(dv, xs_in.p, cell_of.p, cell_cur.p); ++launches;
        size_t tmp_bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cell_cur.p, cell_start.p, (int)(ncells + 1), stream));
        CK(cub_tmp.ensure(tmp_bytes));
        CK(cub::DeviceScan::ExclusiveSum(cub_tmp.p, tmp_bytes, cell_cur.p, cell_start.p, (int)(ncells + 1), stream));
        CK(cudaMemsetAsync(cell_cur.p, 0, (size_t)(ncells + 1) * sizeof(int), stream));
        k_scatter<D><<<blocks_for(n, 256), 256, 0, stream>>>(dv, xs_in.p, cell_of.p, cell_start.p, cell_cur.p, x64.p, x32.p, perm.p); ++launches;
        k_cell_sort<<<blocks_for(ncells, 256), 256, 0, stream>>>(cell_start.p, (int)ncells, perm.p); ++launches;
        k_gather_points<D><<<blocks_for(n, 256), 256, 0, stream>>>(dv, xs_in.p, perm.p, x64.p, x32.p, inv.p); ++launches;
        { int rc = assign_owners(); if (rc) return rc; }
        CK(cudaEventRecord(ev_b, stream));
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_a, ev_b);
        st.ms_build = ms;
        st.ms_upload = 0;                                 // (a periodic margin retry re-records ev_a: no upload in that build)
        if (cudaEventElapsedTime(&ms, ev_a, ev_up) == cudaSuccess && ms > 0 && ms <= st.ms_build) st.ms_upload = ms;
        st.halo_nodes = n_halo;
        return HVB_OK;
    }

    // which rank explores which cell (world > 1): slabs of the sorted order or blocks of the grid, see BlockSpec
    int assign_owners() {
        const int world = std::max(1, prm.world);
        owner_ptr = nullptr;
        if (world == 1) return HVB_OK;
        if (world > 64) { err = "at most 64 ranks"; return HVB_EINVAL; }
        memset(&bspec, 0, sizeof(bspec));
        bspec.world = world; bspec.mode = prm.decomposition ? 1 : 0;
        for (int a = 0; a < 3; ++a) bspec.m[a] = 1;
        if (bspec.mode == 1) {
            // factorise world over the first min(D, 3) axes: every prime factor goes to the axis whose parts are still the
            // thickest in cells (2 x 2 x 2 for 8 ranks in a cube)
            int w = world;
            const int A = std::min(D, 3);
            for (int f = 2; w > 1; ) {
                if (w % f) { ++f; continue; }
                int best = 0; double thick = -1;
                for (int a = 0; a < A; ++a) { const double t = (double)dv.g[a] / bspec.m[a]; if (t > thick) { thick = t; best = a; } }
                bspec.m[best] *= f; w /= f;
            }
            bool ok = true;
            for (int a = 0; a < A; ++a) ok &= (bspec.m[a] <= 16);
            if (!ok) bspec.mode = 0;                       // more than 16 parts along one axis: slabs
        }
        if (bspec.mode == 1) {
            // cuts at the quantiles of the points' coordinates along the cut axes (1024-bin histograms over the bounding box)
            const int A = std::min(D, 3);
            CK(marg.ensure((size_t)A * HVB_CUT_BINS)); CK(h_marg.ensure((size_t)A * HVB_CUT_BINS));
            CK(cudaMemsetAsync(marg.p, 0, (size_t)A * HVB_CUT_BINS * sizeof(unsigned int), stream));
            for (int a = 0; a < A; ++a) {
                if (bspec.m[a] == 1) continue;
                const double width = dv.g[a] * dv.h[a];
                k_coord_hist<D><<<std::min(blocks_for(n, 256), sms * 4), 256, 0, stream>>>(x64.p, (int)n, a, dv.lo[a], 1.0 / width, marg.p + (size_t)a * HVB_CUT_BINS); ++launches;
            }
            CK(cudaMemcpyAsync(h_marg.p, marg.p, (size_t)A * HVB_CUT_BINS * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            for (int a = 0; a < A; ++a) {
                const int m = bspec.m[a];
                if (m == 1) continue;
                const double width = dv.g[a] * dv.h[a];
                const unsigned int* hst = h_marg.p + (size_t)a * HVB_CUT_BINS;
                long long acc = 0; int j = 1;
                for (int c = 0; c < HVB_CUT_BINS && j < m; ++c) {
                    acc += hst[c];
                    while (j < m && acc * m >= (long long)n * j) { bspec.cut[a][j] = dv.lo[a] + width * (double)(c + 1) / HVB_CUT_BINS; ++j; }
                }
                for (; j < m; ++j) bspec.cut[a][j] = dv.lo[a] + width;
            }
        } else {
            for (int k = 0; k <= world; ++k) bspec.bound[k] = n * k / world;       // partition_indices, parallelmesh.jl:52-87
        }
        CK(owner.ensure(n));
        k_assign_owner<D><<<blocks_for(n, 256), 256, 0, stream>>>(dv, bspec, owner.p); ++launches;
        owner_ptr = owner.p;
        return HVB_OK;
    }

    template <int GG>
    void launch_seed_g(const int* seeds, int nseeds, int stride, int cur) {
        k_seed<D, GG><<<std::min(blocks_for((int64_t)nseeds * GG, 128), sms * 16), 128, 0, stream>>>(
            dv, seeds, nseeds, stride, q[cur].p, &sc.p->rnd[cur].qcount, qcap);
        ++launches;
    }
    void launch_seed(const int* seeds, int nseeds, int stride, int cur) {
        launch_seed_g<(D <= 3) ? 4 : 8>(seeds, nseeds, stride, cur);     // descents are few: one tile size per dimension
    }
    template <int GG>
    void launch_expand_g(u32 cnt, int cur, int nxt) {
        k_expand<D, GG><<<std::min(blocks_for((int64_t)cnt * GG, 128), sms * 16), 128, 0, stream>>>(
            dv, q[cur].p, &sc.p->rnd[cur].qcount, &sc.p->rnd[cur].cursor, q[nxt].p, &sc.p->rnd[nxt].qcount, qcap);
    }
    template <int GG>
    int launch_walk_g(const WalkQueue& wq) {
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_walk<D, GG>, 128, 0));
        k_walk<D, GG><<<std::max(1, per_sm) * sms, 128, 0, stream>>>(dv, wq);
        return HVB_OK;
    }
    template <int COOPQ>
    int launch_walk_coop(const WalkQueue& wq) {
        const size_t smem = COOPQ ? 4 * sizeof(CoopShared<D>) : 16;
        CK(cudaFuncSetAttribute(k_walk_coop<D, COOPQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_walk_coop<D, COOPQ>, 128, smem));
        if (debug) fprintf(stderr, "[hvb] k_walk_coop<%d,%d>: %d blocks/SM, %zu B shared per block\n", D, (int)COOPQ, per_sm, smem);
        k_walk_coop<D, COOPQ><<<std::max(1, per_sm) * sms, 128, smem, stream>>>(dv, wq);
        return HVB_OK;
    }
    int launch_walk(const WalkQueue& wq) {
        if (coop && G == 1) return prm.persistent == 3 ? launch_walk_coop<0>(wq) : (prm.persistent == 4 ? launch_walk_coop<2>(wq) : launch_walk_coop<1>(wq));
        switch (G) {
            case 1: return launch_walk_g<1>(wq);
            case 2: return launch_walk_g<2>(wq);
            case 4: return launch_walk_g<4>(wq);
            case 8: return launch_walk_g<8>(wq);
            case 16: return launch_walk_g<16>(wq);
            default: return launch_walk_g<32>(wq);
        }
    }
    void launch_expand(u32 cnt, int cur, int nxt) {
        switch (G) {
            case 1: launch_expand_g<1>(cnt, cur, nxt); break;
            case 2: launch_expand_g<2>(cnt, cur, nxt); break;
            case 4: launch_expand_g<4>(cnt, cur, nxt); break;
            case 8: launch_expand_g<8>(cnt, cur, nxt); break;
            case 16: launch_expand_g<16>(cnt, cur, nxt); break;
            default: launch_expand_g<32>(cnt, cur, nxt); break;
        }
    }

    int alloc_tables(int64_t cap) {
        // a queue entry carries the vertex index in 29 bits (frontier_entry): 5.3e8 records, more than 180 GB of HBM hold
        if (cap >= (1LL << 29)) { err = "vertex capacity beyond 2^29"; return HVB_ENOMEM; }
        vcap = cap;
        CK(vsig.ensure((size_t)cap * (D + 1))); CK(vr.ensure((size_t)cap * D));
        u64 vts = next_pow2((u64)cap * 2);
        u64 ecap = (u64)cap * (D + 1) / 2 + 1024;
        u64 ets = next_pow2(ecap * 2);
        if (ets > (1ULL << 32)) ets = 1ULL << 32;
        CK(vtab.ensure(vts)); CK(etab.ensure(ets));
        dv.vmask = vts - 1; dv.emask = ets - 1;
        qcap = (u32)std::min<u64>(ecap, 0xfffffff0ULL);
        CK(q[0].ensure(qcap)); CK(q[1].ensure(qcap));
        ray_cap = (u32)((P > 0) ? std::max<int64_t>(4096, cap / 16) : cap);
        CK(ray_item.ensure(ray_cap)); CK(ray_u.ensure((size_t)ray_cap * D));
        dv.vsig = vsig.p; dv.vr = vr.p; dv.vcap = (u32)cap; dv.vtab = vtab.p; dv.etab = etab.p;
        dv.ray_item = ray_item.p; dv.ray_u = ray_u.p; dv.ray_cap = ray_cap;
        dv.vcount = &sc.p->vcount; dv.ray_count = &sc.p->ray_count;
        // result buffers and page-locked staging are sized once with the tables (no allocation in steady state)
        for (int i = 0; i < 2; ++i) { CK(out_sig[i].ensure((size_t)cap * (D + 1))); CK(out_r[i].ensure((size_t)cap * D)); }
        CK(key_top.ensure(cap)); CK(key_hi.ensure(cap)); CK(key_lo.ensure(cap)); CK(key_tmp.ensure(cap)); CK(idx[0].ensure(cap)); CK(idx[1].ensure(cap));
        // page-locked staging is pre-sized with the tables only while that stays small (stage() sizes it by the result otherwise)
        if ((size_t)cap * (2 * D + 1) * 8 <= ((size_t)2 << 30)) { CK(h_sig.ensure((size_t)cap * (D + 1))); CK(h_r.ensure((size_t)cap * D)); }
        return HVB_OK;
    }

    // mirrors Scalars + Counters (and optionally one more device word) into page-locked host memory and waits
    int read_scalars(const long long* extra = nullptr, long long* extra_out = nullptr) {
        static_assert(sizeof(Scalars) % 4 == 0 && sizeof(Counters) % 4 == 0, "word copy");
        k_publish<<<1, 64, 0, stream>>>((const u32*)sc.p, (u32*)h_sc.p, (int)(sizeof(Scalars) / 4), (const u32*)ctr.p, (u32*)h_ctr.p,
                                        (int)(sizeof(Counters) / 4), extra, h_extra.p);
        ++launches;
        CK(cudaStreamSynchronize(stream));
        if (extra_out) *extra_out = *h_extra.p;
        return HVB_OK;
    }

    cudaEvent_t pool_event(size_t i) {
        while (ev_pool.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); ev_pool.push_back(e); }
        return ev_pool[i];
    }

    // Periodic domains: the search runs on caller generators + halo and is followed by the certificate; if a ball
    // leaves the pushed planes the margin grows to what the certificate asks for and the search is repeated
    // (the reference repeats its halo step a fixed number of times instead: Create_Discrete_Domain domain.jl:175-213,
    // periodize! :139-166).
    int search(const int64_t* cells, int64_t ncells_in, const int64_t* seed_sig, const double* seed_r, int64_t nseed, int stride) override {
        merged = false; m_nb_built = false;
        // a cloud that turned out to be in non-general position is searched in its perturbed form from then on
        if (perturbed) return resolve_degenerate(cells, ncells_in, nseed);
        int rc = search_general(cells, ncells_in, seed_sig, seed_r, nseed, stride);
        if (rc == HVB_EDEGENERATE && prm.on_degenerate == 2) return resolve_degenerate(cells, ncells_in, nseed);
        return rc;
    }
    int search_general(const int64_t* cells, int64_t ncells_in, const int64_t* seed_sig, const double* seed_r, int64_t nseed, int stride) {
        if (!periodic) return search_once(cells, ncells_in, seed_sig, seed_r, nseed, stride);
        if (nseed > 0) { err = "seed vertices are not supported on a periodic context"; return HVB_EINVAL; }
        if (cells) for (int64_t i = 0; i < ncells_in; ++i) if (cells[i] < 1 || cells[i] > n_user) { err = "Iter names a cell that is not a caller generator"; return HVB_EINVAL; }
        int64_t retries = 0;
        double ms_cert = 0;
        // what the failed attempts cost (their searches, the halo + index rebuilds) stays in the statistics of this call:
        // ms_search and the work counters are sums over all attempts
        hvb_stats_t lost;
        memset(&lost, 0, sizeof(lost));
        for (;;) {
            int rc = search_once(cells, ncells_in, nullptr, nullptr, 0, 0); if (rc) return rc;
            bool ok = false; double need = 0;
            rc = certify(&ok, &need, &ms_cert); if (rc) return rc;
            if (std::max(1, prm.world) > 1) {
                // every rank must build the same halo (its numbering is part of the result): agree on the verdict
                if (ok && !comm) { /* nothing to agree on for this rank; ranks that fail report the missing communicator */ }
                else { double v[2] = {need, ok ? 0.0 : 1.0}; rc = allreduce_max(v, 2); if (rc) return rc; need = v[0]; ok = (v[1] == 0.0); }
            }
            if (ok) break;
            if (++retries > 6) { err = "periodic certificate still fails after 6 margin increases"; return HVB_EINCOMPLETE; }
            lost.ms_search += st.ms_search + st.ms_finalize; lost.ms_expand_kernel += st.ms_expand_kernel; lost.ms_seed += st.ms_seed;
            lost.raycasts += st.raycasts; lost.duplicate_hits += st.duplicate_hits; lost.closed_skips += st.closed_skips;
            lost.candidates_fp32 += st.candidates_fp32; lost.candidates_fp64 += st.candidates_fp64; lost.rows_scanned += st.rows_scanned;
            lost.probe_stages += st.probe_stages; lost.rounds += st.rounds; lost.seeds += st.seeds; lost.kernel_launches += st.kernel_launches;
            lost.expand_launches += st.expand_launches; lost.expand_items += st.expand_items; lost.capacity_retries += st.capacity_retries;
            margin = std::max(1.5 * margin, 1.1 * need);
            if (debug) fprintf(stderr, "[hvb] periodic certificate failed (needs margin %.4g): retry with margin %.4g\n", need, margin);
            CK(cudaEventRecord(ev_a, stream));
            const double build0 = st.ms_build, upload0 = st.ms_upload;
            rc = build_halo(); if (rc) return rc;
            rc = build_index(); if (rc) return rc;
            lost.ms_search += st.ms_build;                 // the rebuild happened inside this hvb_search
            st.ms_build = build0; st.ms_upload = upload0;  // ms_build keeps describing hvb_create / hvb_set_points
        }
        st.periodic_retries = retries;
        st.ms_finalize += ms_cert;
        st.ms_search += lost.ms_search; st.ms_expand_kernel += lost.ms_expand_kernel; st.ms_seed += lost.ms_seed;
        st.raycasts += lost.raycasts; st.duplicate_hits += lost.duplicate_hits; st.closed_skips += lost.closed_skips;
        st.candidates_fp32 += lost.candidates_fp32; st.candidates_fp64 += lost.candidates_fp64; st.rows_scanned += lost.rows_scanned;
        st.probe_stages += lost.probe_stages; st.rounds += lost.rounds; st.seeds += lost.seeds; st.kernel_launches += lost.kernel_launches;
        st.expand_launches += lost.expand_launches; st.expand_items += lost.expand_items; st.capacity_retries += lost.capacity_retries;
        // the certified margin, in units of the generator spacing, is kept for the next cloud on this context
        if (retries > 0 && margin_spacing > 0) margin_rel_certified = std::max(margin_rel_certified, margin / margin_spacing);
        return HVB_OK;
    }

    int certify(bool* ok, double* need, double* ms_total) {
        CK(cudaEventRecord(ev_p0, stream));
        CK(vflags.ensure(std::max<int64_t>(nvert, 1)));
        CK(cudaMemsetAsync(cert.p, 0, sizeof(CertOut), stream));
        if (nvert > 0) {
            k_certify<D><<<blocks_for(nvert, 128), 128, 0, stream>>>(out_sig[res].p, out_r[res].p, (u32)nvert, (long long)n_user, (long long)n,
                                                                    xs_in.p, halo_origin.p, planes.p, pcert, vflags.p, cert.p);
            ++launches;
        }
        static_assert(sizeof(CertOut) % 4 == 0, "word copy");
        k_publish<<<1, 64, 0, stream>>>((const u32*)cert.p, (u32*)h_cert.p, (int)(sizeof(CertOut) / 4), nullptr, nullptr, 0, nullptr, nullptr);
        ++launches;
        CK(cudaEventRecord(ev_p1, stream));
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_p0, ev_p1);
        *ms_total += ms;
        double excess;
        memcpy(&excess, &h_cert.p->max_excess_bits, sizeof(double));
        *need = excess;
        if (h_cert.p->self_neighbor > 0) {
            err = "a periodic cell neighbours its own image: too few generators for this periodic domain";
            return HVB_EINCOMPLETE;
        }
        *ok = (h_cert.p->on_pushed_plane == 0) && (excess <= margin);
        if (h_cert.p->on_pushed_plane > 0) *need = std::max(*need, 1.5 * margin);
        st.unique_vertices = h_cert.p->canonical;
        st.kernel_launches = launches;
        have_flags = *ok;
        return HVB_OK;
    }

    int halo_count(int64_t* nhalo, int32_t* npairs, double* mg) override {
        if (nhalo) *nhalo = n_halo;
        if (npairs) *npairs = halo.npairs;
        if (mg) *mg = periodic ? margin : 0.0;
        return HVB_OK;
    }
    int fetch_halo(int64_t* origin, int32_t* mult, double* xs) override {
        if (!periodic || n_halo == 0) return HVB_OK;
        CK(cudaSetDevice(prm.device));
        std::vector<int> o(n_halo);
        std::vector<signed char> m((size_t)n_halo * halo.npairs);
        CK(cudaMemcpyAsync(o.data(), halo_origin.p, (size_t)n_halo * sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(m.data(), halo_mult.p, (size_t)n_halo * halo.npairs, cudaMemcpyDeviceToHost, stream));
        if (xs) CK(cudaMemcpyAsync(xs, xs_in.p + (size_t)n_user * D, (size_t)n_halo * D * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (origin) for (int64_t i = 0; i < n_halo; ++i) origin[i] = (int64_t)o[i] + 1;
        if (mult) for (size_t i = 0; i < m.size(); ++i) mult[i] = (int32_t)m[i];
        return HVB_OK;
    }
    int fetch_vertex_flags(uint8_t* flags) override {
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        if (!flags) { err = "null output"; return HVB_EINVAL; }
        CK(cudaSetDevice(prm.device));
        if (!periodic || !have_flags) { memset(flags, 1, (size_t)(merged ? m_nvert : nvert)); return HVB_OK; }   // without a halo every row is its own representative
        if (nvert > 0) CK(cudaMemcpyAsync(flags, vflags.p, (size_t)nvert, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return HVB_OK;
    }

    // ---- non-general position: resolved by perturbation + merge ----------------------------------------
    // The reference enumerates the edges of a vertex with more than d + 1 cospherical generators by linear algebra on the
    // cone of the vertex (FastEdgeIterator, edgeiterate.jl:82-780) and returns ONE vertex whose signature lists all of them
    // (raycast.jl:870-969).  A cone enumeration is branchy, variable-length, sequential work -- the opposite of what the
    // walk kernel is good at.  This backend gets the same RESULT from the general-position machinery:
    //   1. the generators are moved by a deterministic pseudo-random offset of relative size HVB_PERTURB_REL (an explicit
    //      simulation of simplicity): the perturbed cloud is in general position and its Delaunay triangulation restricted
    //      to a cospherical set S is a triangulation of conv(S);
    //   2. the search runs on the perturbed cloud unchanged;
    //   3. coordinates are solved from the CALLER's generators (canonical_vertex on dv.xcan): every simplex of the
    //      triangulation of conv(S) gets the circumcentre of S, identical up to rounding; simplices whose d + 1 generators
    //      lie in one hyperplane of the caller's cloud are slivers between two cospherical cells and are dropped (k_final_rows);
    //   4. rows with equal coordinates are merged (host, rare path): the union of their signatures is S.
    // Checked against Qhull's Voronoi diagram of the same cloud (which merges cospherical facets), tests/test_gpu_degenerate.py.
    DBuf<double> xs_orig, xcan;
    bool perturbed = false;              // xs_in / x64 / x32 hold perturbed generators, xs_orig / xcan the caller's
    bool merged = false;                 // the result of the last search are the variable-length rows below
    bool m_nb_built = false;
    std::vector<int64_t> m_off, m_ids, m_nb_off, m_nb_ids;
    std::vector<double> m_r;
    int64_t m_nvert = 0, m_maxlen = 0, m_degenerate = 0;
    bool report_degenerate() const { return prm.on_degenerate == 0 || (prm.on_degenerate == 2 && !perturbed); }

    int resolve_degenerate(const int64_t* cells, int64_t ncells_in, int64_t nseed) {
        // unbounded domains: the hull of such a cloud (the faces of a lattice) has coplanar generators whose perturbed
        // simplices have balls of arbitrary size; telling a generator ON such a ball from one inside it is beyond FP64
        // Iter subsets: the simplices of a cospherical set that touch none of the explored cells are not found, their
        // generators would be missing from the merged signature
        if (periodic || std::max(1, prm.world) > 1 || nseed > 0 || P == 0 || cells != nullptr) {
            err = "non-general position: a vertex with more than dim+1 cospherical generators was met (resolving it is available on bounded, non-periodic domains, one GPU, searches over all cells without seed vertices)";
            return HVB_EDEGENERATE;
        }
        CK(cudaSetDevice(prm.device));
        if (!perturbed) {
            const double build0 = st.ms_build, upload0 = st.ms_upload;
            CK(xs_orig.ensure((size_t)n * D));
            CK(cudaMemcpyAsync(xs_orig.p, xs_in.p, (size_t)n * D * sizeof(double), cudaMemcpyDeviceToDevice, stream));
            k_perturb<<<blocks_for(n * D, 256), 256, 0, stream>>>(xs_orig.p, xs_in.p, (size_t)n * D, D, HVB_PERTURB_REL * dv.ext); ++launches;
            CK(cudaEventRecord(ev_a, stream));
            int rc = check_points(xs_in.p, n, nullptr);
            if (rc) {
                CK(cudaMemcpyAsync(xs_in.p, xs_orig.p, (size_t)n * D * sizeof(double), cudaMemcpyDeviceToDevice, stream));
                CK(cudaStreamSynchronize(stream));
                err = "non-general position: a generator lies too close to the boundary for the perturbation that resolves it";
                return HVB_EDEGENERATE;
            }
            rc = build_index(); if (rc) return rc;
            CK(xcan.ensure((size_t)n * D));
            k_gather_canon<D><<<blocks_for(n, 256), 256, 0, stream>>>(xs_orig.p, perm.p, (int)n, xcan.p); ++launches;
            dv.xcan = xcan.p;
            dv.t_min = HVB_TMIN_REL * dv.ext;
            perturbed = true;
            st.ms_build = build0; st.ms_upload = upload0;
        }
        const int nb_save = prm.neighbors;
        prm.neighbors = 0;                       // the lists of a merged mesh are built from the merged rows (build_merged_neighbors)
        int rc = search_once(cells, ncells_in, nullptr, nullptr, 0, 0);
        prm.neighbors = nb_save;
        if (rc) return rc;
        return merge_result();
    }

    // step 4: rows with equal coordinates -> one vertex with the union of the signatures (hvb_nongeneral.hpp)
    int merge_result() {
        const int64_t* sig; const double* r; int64_t nrow;
        int rc = view_vertices(&sig, &r, &nrow); if (rc) return rc;
        merge_rows(D, nrow, sig, r, HVB_MERGE_REL * dv.ext, prm.sort_output != 0, m_off, m_ids, m_r, m_maxlen, m_degenerate);
        m_nvert = (int64_t)m_off.size() - 1;
        merged = true; m_nb_built = false;
        st.vertices = m_nvert; st.unique_vertices = m_nvert; st.degenerate = m_degenerate;
        return HVB_OK;
    }

    // neighbour lists of a merged mesh: cells that share a FULL interface (merged_neighbors, hvb_nongeneral.hpp)
    int build_merged_neighbors() {
        if (m_nb_built) return HVB_OK;
        std::vector<int64_t> redge; std::vector<double> rdir;
        if (nrays > 0) {
            redge.resize((size_t)nrays * D); rdir.resize((size_t)nrays * D);
            int rc = fetch_rays(redge.data(), nullptr, rdir.data(), nullptr); if (rc) return rc;
        }
        merged_neighbors(D, n, m_nvert, m_off.data(), m_ids.data(), m_r.data(), nrays, redge.data(), rdir.data(), m_nb_off, m_nb_ids);
        m_nb_built = true;
        return HVB_OK;
    }

    // variable-length rows (hvb_fetch_vertices_var): works for every result; general position gives off[v] = v (dim + 1)
    int fetch_vertices_var(int64_t* off, int64_t* ids, double* r) override {
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        if (merged) {
            if (off) memcpy(off, m_off.data(), m_off.size() * sizeof(int64_t));
            if (ids && !m_ids.empty()) memcpy(ids, m_ids.data(), m_ids.size() * sizeof(int64_t));
            if (r && !m_r.empty()) memcpy(r, m_r.data(), m_r.size() * sizeof(double));
            return HVB_OK;
        }
        if (off) for (int64_t v = 0; v <= nvert; ++v) off[v] = v * (D + 1);
        return fetch_vertices(ids, r);
    }

    // ---- convex hull by the facet walk (hvb_hull.cuh; replaces systematic_chull, chull.jl:241-387) -------------------
    DBuf<int> h_fsig; DBuf<u32> h_fitem; DBuf<double> h_fu; DBuf<u64> h_ftab, h_rtab; DBuf<unsigned long long> h_arg;
    // ---- convex hull by gift wrapping (hvb_wrap.cuh): the default ----------------------------------------------------
    DBuf<unsigned char> w_wq; DBuf<double> w_pc1, w_pc2; DBuf<int> w_pid; DBuf<u64> w_q[2]; DBuf<u32> w_words; DBuf<WrapSeed> w_seed;
    HBuf<u32> hw_words;
    int convex_hull(int method) override {
        if (method == 1) return convex_hull_walk();
        if (method != 0) { err = "hvb_convex_hull_via: method 0 (gift wrapping) or 1 (walk around the unbounded 2-faces)"; return HVB_EINVAL; }
        if (P != 0 || periodic) { err = "the convex hull is computed on the unbounded domain: create the context without planes"; return HVB_EINVAL; }
        if (std::max(1, prm.world) > 1) { err = "the convex hull runs on one GPU"; return HVB_EINVAL; }
        if (n <= D) { err = "the convex hull needs more than dim generators"; return HVB_EINVAL; }
        CK(cudaSetDevice(prm.device));
        have_result = false; staged = false; staged_r = staged_sig64 = staged_sig32 = false; nb_staged32 = false; nb_total = -1; have_flags = false;
        launches = 0;
        CK(h_arg.ensure(16)); CK(w_words.ensure(32)); CK(hw_words.ensure(32)); CK(w_seed.ensure(16));
        const int tb = sms * 4;
        const int grid_small = sms * 2;
        u32 fcap = (u32)std::min<int64_t>(std::max<int64_t>(prm.vertex_capacity > 0 ? prm.vertex_capacity : 0, 1 << 16), 0x07ffffff);
        int64_t rounds = 0;
        u32 nf = 0;
        st.capacity_retries = 0;
        for (int attempt = 0;; ++attempt) {
            HullDev<D> hd;
            WrapDev<D> wd;
            const u64 fts = next_pow2(2 * (u64)fcap), rts = next_pow2(2 * (u64)fcap * D);
            const u32 wq_cap = fcap, wqueue = (u32)std::min<u64>((u64)fcap * D, 0x7fffffffULL);
            const u32 pcap = std::max<u32>(wq_cap, (u32)tb * 128u);
            CK(h_fsig.ensure((size_t)fcap * D)); CK(h_fitem.ensure(fcap)); CK(h_fu.ensure((size_t)fcap * D)); CK(h_ftab.ensure(fts)); CK(h_rtab.ensure(rts));
            CK(w_wq.ensure((size_t)wq_cap * sizeof(WrapQuery<D>))); CK(w_pc1.ensure(pcap)); CK(w_pc2.ensure(pcap)); CK(w_pid.ensure(pcap));
            CK(w_q[0].ensure(wqueue)); CK(w_q[1].ensure(wqueue));
            hd.fsig = h_fsig.p; hd.fitem = h_fitem.p; hd.fu = h_fu.p; hd.fcount = &sc.p->ray_count; hd.fcap = fcap;
            hd.ftab = h_ftab.p; hd.fmask = fts - 1; hd.rtab = h_rtab.p; hd.rmask = rts - 1;
            wd.wq = (WrapQuery<D>*)w_wq.p; wd.wq_cap = wq_cap; wd.pc1 = w_pc1.p; wd.pc2 = w_pc2.p; wd.pid = w_pid.p; wd.pcap = pcap;
            wd.q[0] = w_q[0].p; wd.q[1] = w_q[1].p; wd.qcap = wqueue; wd.qcount = w_words.p; wd.nq = w_words.p + 2; wd.seed = w_seed.p;
            wrap_tolerances<D>(dv.ext, wd.E32, wd.tinyA);
            if (!prm.fp32_filter) wd.E32 = INFINITY;
            CK(cudaEventRecord(ev_a, stream));            // device time of the walk: buffers exist (a warm context allocates nothing)
            CK(cudaMemsetAsync(h_ftab.p, 0, fts * sizeof(u64), stream));
            CK(cudaMemsetAsync(h_rtab.p, 0, rts * sizeof(u64), stream));
            CK(cudaMemsetAsync(ctr.p, 0, sizeof(Counters), stream));
            CK(cudaMemsetAsync(sc.p, 0, sizeof(Scalars), stream));
            CK(cudaMemsetAsync(h_arg.p, 0, 16 * sizeof(unsigned long long), stream));
            CK(cudaMemsetAsync(w_words.p, 0xff, 32 * sizeof(u32), stream));
            k_wrap_extremes<D><<<std::min(blocks_for(n, 256), sms * 4), 256, 0, stream>>>(x64.p, (int)n, h_arg.p, 0, w_words.p + 8);
            k_wrap_extremes<D><<<std::min(blocks_for(n, 256), sms * 4), 256, 0, stream>>>(x64.p, (int)n, h_arg.p, 1, w_words.p + 8);
            k_wrap_init<D><<<1, 1, 0, stream>>>(wd, w_words.p + 8);
            launches += 3;
            int cur = 0;
            rounds = 0;
            bool overflow = false;
            // the host enqueues `batch` rounds and only then looks at the queue: a round after the end is three empty launches
            for (int batch = D + 3;; batch = 4) {
                for (int b = 0; b < batch; ++b) {
                    k_wrap_prepare<D><<<grid_small, 128, 0, stream>>>(dv, hd, wd, cur);
                    k_wrap_scan<D><<<tb, 128, 0, stream>>>(dv, wd, cur, tb);
                    k_wrap_commit<D><<<grid_small, 128, 0, stream>>>(dv, hd, wd, cur, tb);
                    launches += 3; ++rounds;
                    cur = 1 - cur;
                }
                CK(cudaMemcpyAsync(hw_words.p, w_words.p, 4 * sizeof(u32), cudaMemcpyDeviceToHost, stream));
                int rc = read_scalars(); if (rc) return rc;
                if (h_ctr.p->flags & FLAG_OVERFLOW_MASK) { overflow = true; break; }
                if ((h_ctr.p->degenerate > 0 || (h_ctr.p->flags & FLAG_DEGEN)) && prm.on_degenerate != 1) {
                    st.degenerate = (int64_t)std::max<u64>(h_ctr.p->degenerate, 1);
                    err = "non-general position: a hull facet with more than dim generators was met";
                    return HVB_EDEGENERATE;
                }
                if (hw_words.p[cur] == 0) break;
                if (rounds > 1000000) { err = "hull walk does not terminate"; return HVB_EINCOMPLETE; }
            }
            CK(cudaEventRecord(ev_c, stream));
            if (overflow) {
                if (attempt >= 4 || fcap >= 0x07ffffff) { err = "hull walk: capacity exhausted (raise vertex_capacity)"; return HVB_ENOMEM; }
                fcap = (u32)std::min<u64>((u64)fcap * 8, 0x07ffffff);
                ++st.capacity_retries;
                continue;
            }
            if (h_ctr.p->seed_fail > 0) { err = "hull walk: a query found no generator (fewer than dim + 1 generators in general position?)"; return HVB_EINCOMPLETE; }
            nf = std::min<u32>(h_sc.p->ray_count, fcap);
            nvert = 0; nrays = 0;
            if (nf > 0) { CK(ray_edge.ensure((size_t)nf * D)); CK(ray_base.ensure((size_t)nf * D)); CK(ray_dir.ensure((size_t)nf * D)); CK(ray_node.ensure(nf)); }
            CK(cudaEventRecord(ev_b, stream));
            if (nf > 0) {
                k_final_facets_wrap<D><<<blocks_for(nf, 128), 128, 0, stream>>>(dv, hd, perm.p, nf, ray_edge.p, ray_base.p, ray_dir.p, ray_node.p, &sc.p->ray_out, &sc.p->pflags); ++launches;
            }
            break;
        }
        CK(cudaEventRecord(ev_d, stream));
        int rc = read_scalars(); if (rc) return rc;
        nrays = h_sc.p->ray_out;
        if (h_sc.p->pflags > 0 && prm.on_degenerate != 1) { err = "non-general position: a hull facet whose generators do not span a hyperplane"; return HVB_EDEGENERATE; }
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_a, ev_c); st.ms_search = ms;
        cudaEventElapsedTime(&ms, ev_b, ev_d); st.ms_finalize = ms;
        const Counters& c = *h_ctr.p;
        st.ms_expand_kernel = st.ms_search; st.expand_launches = rounds; st.expand_items = (int64_t)c.raycasts; st.rounds = rounds;
        st.vertices = 0; st.unique_vertices = 0; st.rays = nrays; st.raycasts = (int64_t)c.raycasts; st.duplicate_hits = (int64_t)c.dup_hits; st.closed_skips = (int64_t)c.closed_skips;
        st.candidates_fp32 = (int64_t)c.cand32; st.candidates_fp64 = (int64_t)c.cand64; st.rows_scanned = 0; st.probe_stages = (int64_t)c.stages;
        st.seeds = 1; st.degenerate = (int64_t)c.degenerate; st.kernel_launches = launches; st.ms_seed = 0; st.ms_neighbors = 0;
        st.ms_rows_sort = 0; st.ms_stage_wait = 0; st.rejected = 0; st.suboptimal = 0; st.exchange_bytes = 0; st.periodic_retries = 0;
        have_result = true;
        return HVB_OK;
    }
    // the first device hull walk (hvb_hull.cuh): around the unbounded 2-faces of the diagram with min-t queries.  Kept as
    // hvb_convex_hull_via(ctx, 1): same facets, a cross-check of the wrapping on the Voronoi side
    int convex_hull_walk() {
        if (P != 0 || periodic) { err = "the convex hull is computed on the unbounded domain: create the context without planes"; return HVB_EINVAL; }
        if (std::max(1, prm.world) > 1) { err = "the convex hull runs on one GPU"; return HVB_EINVAL; }
        CK(cudaSetDevice(prm.device));
        have_result = false; staged = false; staged_r = staged_sig64 = staged_sig32 = false; nb_staged32 = false; nb_total = -1; have_flags = false;
        launches = 0;
        int64_t cap = std::max<int64_t>(vcap, prm.vertex_capacity > 0 ? prm.vertex_capacity : estimate_vertices(D, n, 0));
        CK(cudaEventRecord(ev_a, stream));
        if (cap != vcap || !vsig.p) { int rc = alloc_tables(cap); if (rc) return rc; }
        HullDev<D> hd;
        const u32 fcap = (u32)std::min<int64_t>(vcap, 0x7fffffff);
        const u64 fts = next_pow2(2 * (u64)fcap), rts = next_pow2(2 * (u64)fcap * D);
        CK(h_fsig.ensure((size_t)fcap * D)); CK(h_fitem.ensure(fcap)); CK(h_fu.ensure((size_t)fcap * D)); CK(h_ftab.ensure(fts)); CK(h_rtab.ensure(rts)); CK(h_arg.ensure(1));
        hd.fsig = h_fsig.p; hd.fitem = h_fitem.p; hd.fu = h_fu.p; hd.fcount = &sc.p->ray_count; hd.fcap = fcap;
        hd.ftab = h_ftab.p; hd.fmask = fts - 1; hd.rtab = h_rtab.p; hd.rmask = rts - 1;
        CK(cudaMemsetAsync(h_ftab.p, 0, fts * sizeof(u64), stream));
        CK(cudaMemsetAsync(h_rtab.p, 0, rts * sizeof(u64), stream));
        CK(cudaMemsetAsync(ctr.p, 0, sizeof(Counters), stream));
        CK(cudaMemsetAsync(sc.p, 0, sizeof(Scalars), stream));
        CK(cudaMemsetAsync(h_arg.p, 0, sizeof(unsigned long long), stream));
        CK(cudaMemsetAsync(active.p, 1, n, stream));
        const int axis = 0;
        k_argmax_axis<D><<<std::min(blocks_for(n, 256), sms * 4), 256, 0, stream>>>(x64.p, (int)n, axis, h_arg.p); ++launches;
        CK(cudaMemcpyAsync(h_extra.p, h_arg.p, sizeof(long long), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        const int start = (int)((unsigned long long)*h_extra.p & 0xffffffffULL);
        int cur = 0;
        // most walks near the hull that find nothing in the first probe ball find nothing at all: the second stage is the
        // half-space (exact; a candidate met on the way shrinks it to a ball again)
        Dev<D> dh = dv;
        dh.probe_growth = 1e9;
        k_hull_seed<D><<<1, 32, 0, stream>>>(dh, hd, start, axis, q[cur].p, &sc.p->rnd[cur].qcount, qcap); ++launches;
        int64_t rounds = 0, items = 0;
        size_t n_ev = 0;
        for (;;) {
            int rc = read_scalars(); if (rc) return rc;
            if (h_ctr.p->flags & FLAG_OVERFLOW_MASK) { err = "hull walk: capacity exhausted (raise vertex_capacity)"; return HVB_ENOMEM; }
            if ((h_ctr.p->degenerate > 0 || (h_ctr.p->flags & FLAG_DEGEN)) && prm.on_degenerate != 1) {
                st.degenerate = (int64_t)std::max<u64>(h_ctr.p->degenerate, 1);
                err = "non-general position: a hull facet with more than dim generators was met";
                return HVB_EDEGENERATE;
            }
            const u32 cnt = h_sc.p->rnd[cur].qcount;
            if (cnt == 0) break;
            const int nxt = 1 - cur;
            CK(cudaMemsetAsync(&sc.p->rnd[nxt], 0, sizeof(Round), stream));
            cudaEvent_t e0 = pool_event(n_ev++), e1 = pool_event(n_ev++);
            CK(cudaEventRecord(e0, stream));
            k_hull_expand<D><<<std::min(blocks_for((int64_t)cnt * 32, 128), sms * 16), 128, 0, stream>>>(dh, hd, q[cur].p, &sc.p->rnd[cur].qcount, &sc.p->rnd[cur].cursor,
                                                                                       q[nxt].p, &sc.p->rnd[nxt].qcount, qcap);
            CK(cudaEventRecord(e1, stream));
            ++launches; ++rounds; items += cnt;
            cur = nxt;
            if (rounds > 100000) { err = "hull walk does not terminate"; return HVB_EINCOMPLETE; }
        }
        if (h_ctr.p->seed_fail > 0 && h_sc.p->ray_count == 0) { err = "hull walk: no first facet found"; return HVB_EINCOMPLETE; }
        CK(cudaEventRecord(ev_b, stream));
        const u32 nf = std::min<u32>(h_sc.p->ray_count, fcap);
        nvert = 0; nrays = 0;
        if (nf > 0) {
            CK(ray_edge.ensure((size_t)nf * D)); CK(ray_base.ensure((size_t)nf * D)); CK(ray_dir.ensure((size_t)nf * D)); CK(ray_node.ensure(nf));
            k_final_facets<D><<<blocks_for(nf, 128), 128, 0, stream>>>(dh, hd, perm.p, nf, ray_edge.p, ray_base.p, ray_dir.p, ray_node.p, &sc.p->ray_out); ++launches;
        }
        CK(cudaEventRecord(ev_d, stream));
        int rc = read_scalars(); if (rc) return rc;
        nrays = h_sc.p->ray_out;
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_a, ev_b); st.ms_search = ms;
        cudaEventElapsedTime(&ms, ev_b, ev_d); st.ms_finalize = ms;
        double kms = 0;
        for (size_t i = 0; i + 1 < n_ev; i += 2) { cudaEventElapsedTime(&ms, ev_pool[i], ev_pool[i + 1]); kms += ms; }
        const Counters& c = *h_ctr.p;
        st.ms_expand_kernel = kms; st.expand_launches = rounds; st.expand_items = items; st.rounds = rounds;
        st.vertices = 0; st.unique_vertices = 0; st.rays = nrays; st.raycasts = (int64_t)c.raycasts; st.duplicate_hits = 0; st.closed_skips = 0;
        st.candidates_fp32 = (int64_t)c.cand32; st.candidates_fp64 = (int64_t)c.cand64; st.rows_scanned = (int64_t)c.rows; st.probe_stages = (int64_t)c.stages;
        st.seeds = 1; st.degenerate = (int64_t)c.degenerate; st.kernel_launches = launches; st.capacity_retries = 0; st.ms_seed = 0; st.ms_neighbors = 0;
        st.ms_rows_sort = 0; st.ms_stage_wait = 0; st.rejected = 0; st.suboptimal = 0; st.exchange_bytes = 0; st.periodic_retries = 0;
        have_result = true;
        return HVB_OK;
    }

    // which caller cells this context owns (multi-GPU slabs; all of them otherwise)
    int fetch_owned(uint8_t* owned) override {
        if (!owned) { err = "null output"; return HVB_EINVAL; }
        const int64_t n_list = periodic ? n_user : n;
        const int world = std::max(1, prm.world), rank = std::min(std::max(0, prm.rank), world - 1);
        if (world == 1) { memset(owned, 1, (size_t)n_list); return HVB_OK; }
        CK(cudaSetDevice(prm.device));
        CK(own_mask.ensure(n));
        k_own_mask<<<blocks_for(n, 256), 256, 0, stream>>>(perm.p, (int)n, owner_ptr, rank, own_mask.p); ++launches;
        CK(cudaMemcpyAsync(owned, own_mask.p, (size_t)n_list, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return HVB_OK;
    }

    // volumes of the cells of the caller's generators from the current result rows (hvb_geometry.cuh)
    DBuf<long long> vol_acc;
    DBuf<double> vol_dev;
    int cell_volumes(double* vol) override {
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        if (!vol) { err = "null output"; return HVB_EINVAL; }
        if (seed_prefix > 0) { err = "cell volumes need all vertices of the cells: not available after a search with seed vertices"; return HVB_ESTATE; }
        CK(cudaSetDevice(prm.device));
        const long long n_list = periodic ? n_user : n;
        CK(vol_acc.ensure(n_list)); CK(vol_dev.ensure(n_list)); CK(vol_sat.ensure(n_list));
        CK(cudaMemsetAsync(vol_acc.p, 0, (size_t)n_list * sizeof(long long), stream));
        CK(cudaMemsetAsync(vol_sat.p, 0, (size_t)n_list, stream));
        double fact = 1.0;
        for (int k = 2; k <= D; ++k) fact *= k;
        // fixed point: 2^52 units per ext^D (the bounding box volume is at most ext^D); 1/d! is folded into the scale so
        // that the accumulators hold volumes, with 11 bits of headroom for partial sums of either sign
        const double scale = ldexp(1.0, 52) / (pow(dv.ext, (double)D) * fact);
        if (perturbed) {
            // a mesh resolved from non-general position: the volumes of the PERTURBED cells (the perturbed diagram is simple,
            // which is what the flag formula needs; they differ from the caller's by O(HVB_PERTURB_REL)), from every vertex
            // record of the walk -- the slivers the result rows leave out are vertices of that diagram too
            const u32 nrec = std::min<u32>(h_sc.p->vcount, (u32)vcap);
            long long* rows = out_sig[1 - res].p;
            CK(cudaMemsetAsync(&sc.p->pad2, 0, sizeof(u32), stream));
            k_rows_from_records<D><<<blocks_for(nrec, 128), 128, 0, stream>>>(dv, perm.p, nrec, rows, &sc.p->pad2); ++launches;
            CK(cudaMemcpyAsync(h_extra.p, &sc.p->pad2, sizeof(u32), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            const u32 cnt = *(const u32*)h_extra.p;
            if (cnt > 0) { k_cell_volumes<D><<<blocks_for(cnt, 128), 128, 0, stream>>>(rows, cnt, xs_in.p, (long long)n, n_list, planes.p, scale, vol_acc.p, vol_sat.p); ++launches; }
        } else
        if (nvert > 0) {
            k_cell_volumes<D><<<blocks_for(nvert, 128), 128, 0, stream>>>(out_sig[res].p, (u32)nvert, xs_in.p, (long long)n, n_list, planes.p, scale, vol_acc.p, vol_sat.p);
            ++launches;
        }
        k_volumes_finish<<<blocks_for(n_list, 256), 256, 0, stream>>>(vol_acc.p, 1.0 / (scale * fact), vol_dev.p, n_list, vol_sat.p); ++launches;
        if (nrays > 0) { k_volumes_unbounded<<<blocks_for(nrays * D, 256), 256, 0, stream>>>(ray_edge.p, (long long)nrays * D, n_list, vol_dev.p); ++launches; }
        CK(cudaMemcpyAsync(vol, vol_dev.p, (size_t)n_list * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        st.kernel_launches = launches;
        return HVB_OK;
    }

    // integrals of 1, x_a and x_a x_b over every cell (hvb_cell_moments): same rows, same completeness rule, same fixed-point
    // accumulation as the volumes
    DBuf<double> mom_dev;
    DBuf<unsigned char> vol_sat;           // cells (or list entries) one of whose terms left the fixed-point range: they get NaN
    int cell_moments(double* vol, double* first, double* second) override {
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        if (seed_prefix > 0) { err = "cell moments need all vertices of the cells: not available after a search with seed vertices"; return HVB_ESTATE; }
        if (std::max(1, prm.world) > 1) { err = "hvb_cell_moments runs on a single-GPU context"; return HVB_EINVAL; }
        CK(cudaSetDevice(prm.device));
        const int NM = 1 + D + D * (D + 1) / 2;
        const long long n_list = periodic ? n_user : n;
        CK(vol_acc.ensure((size_t)n_list * NM)); CK(mom_dev.ensure((size_t)n_list * (1 + D + D * D))); CK(vol_sat.ensure(n_list));
        CK(cudaMemsetAsync(vol_acc.p, 0, (size_t)n_list * NM * sizeof(long long), stream));
        CK(cudaMemsetAsync(vol_sat.p, 0, (size_t)n_list, stream));
        double fact = 1.0;
        for (int k = 2; k <= D; ++k) fact *= k;
        // one fixed-point scale per degree: 2^52 units per ext^(D + degree) (local coordinates are bounded by the cell's diameter)
        const double s0 = ldexp(1.0, 52) / (pow(dv.ext, (double)D) * fact), s1 = s0 / dv.ext, s2 = s1 / dv.ext;
        const long long* rows = out_sig[res].p;
        u32 cnt = (u32)nvert;
        if (perturbed) {
            const u32 nrec = std::min<u32>(h_sc.p->vcount, (u32)vcap);
            long long* tmp = out_sig[1 - res].p;
            CK(cudaMemsetAsync(&sc.p->pad2, 0, sizeof(u32), stream));
            k_rows_from_records<D><<<blocks_for(nrec, 128), 128, 0, stream>>>(dv, perm.p, nrec, tmp, &sc.p->pad2); ++launches;
            CK(cudaMemcpyAsync(h_extra.p, &sc.p->pad2, sizeof(u32), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            cnt = *(const u32*)h_extra.p; rows = tmp;
        }
        if (cnt > 0) { k_cell_moments<D><<<blocks_for(cnt, 128), 128, 0, stream>>>(rows, cnt, xs_in.p, (long long)n, n_list, planes.p, s0, s1, s2, vol_acc.p, vol_sat.p); ++launches; }
        double* dvol = mom_dev.p; double* dfirst = mom_dev.p + n_list; double* dsecond = mom_dev.p + n_list * (1 + D);
        k_moments_finish<D><<<blocks_for(n_list, 128), 128, 0, stream>>>(vol_acc.p, xs_in.p, n_list, 1.0 / (s0 * fact), 1.0 / (s1 * fact), 1.0 / (s2 * fact), dvol, dfirst, dsecond, vol_sat.p); ++launches;
        if (nrays > 0) { k_moments_unbounded<<<blocks_for(nrays * D, 256), 256, 0, stream>>>(ray_edge.p, (long long)nrays * D, n_list, D, dvol, dfirst, dsecond); ++launches; }
        if (vol) CK(cudaMemcpyAsync(vol, dvol, (size_t)n_list * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (first) CK(cudaMemcpyAsync(first, dfirst, (size_t)n_list * D * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (second) CK(cudaMemcpyAsync(second, dsecond, (size_t)n_list * D * D * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        st.kernel_launches = launches;
        return HVB_OK;
    }

    // clean_affected! (meshrefine.jl:126-149): which vertices of the caller's old mesh survive the new generators
    DBuf<unsigned char> ca_keep, ca_aff;
    int clean_affected(const int64_t* sig, const double* r, int64_t nv, int stride, int64_t first_new, int64_t n_new, uint8_t* keep, uint8_t* affected) override {
        if (merged || perturbed) { err = "hvb_clean_affected: not available on a context whose cloud was resolved from non-general position"; return HVB_ESTATE; }
        if (periodic) { err = "refinement is not supported on a periodic context"; return HVB_EINVAL; }
        if (nv < 0 || stride < 1 || (nv > 0 && (!sig || !r || !keep)) || !affected) { err = "bad arguments"; return HVB_EINVAL; }
        if (first_new < 1 || n_new < 0 || first_new + n_new - 1 > n) { err = "the new generators must be a range of the context's ids"; return HVB_EINVAL; }
        CK(cudaSetDevice(prm.device));
        CK(seed_sig_dev.ensure((size_t)std::max<int64_t>(nv, 1) * stride)); CK(seed_r_dev.ensure((size_t)std::max<int64_t>(nv, 1) * D));
        CK(ca_keep.ensure(std::max<int64_t>(nv, 1))); CK(ca_aff.ensure(n));
        CK(cudaMemsetAsync(ca_aff.p, 0, n, stream));
        CK(cudaMemsetAsync(&sc.p->pflags, 0, sizeof(u32), stream));
        if (nv > 0) {
            CK(cudaMemcpyAsync(seed_sig_dev.p, sig, (size_t)nv * stride * 8, cudaMemcpyHostToDevice, stream));
            CK(cudaMemcpyAsync(seed_r_dev.p, r, (size_t)nv * D * 8, cudaMemcpyHostToDevice, stream));
            k_clean_affected<D><<<blocks_for(nv, 128), 128, 0, stream>>>(dv, perm.p, seed_sig_dev.p, seed_r_dev.p, nv, stride, xs_in.p,
                                                                        first_new - 1, first_new - 1 + n_new, ca_keep.p, ca_aff.p, &sc.p->pflags);
            ++launches;
            CK(cudaMemcpyAsync(keep, ca_keep.p, (size_t)nv, cudaMemcpyDeviceToHost, stream));
        }
        CK(cudaMemcpyAsync(affected, ca_aff.p, (size_t)n, cudaMemcpyDeviceToHost, stream));
        int rc = read_scalars(); if (rc) return rc;
        CK(cudaGetLastError());
        if (h_sc.p->pflags) { err = "a vertex row without a generator id of this context"; return HVB_EINVAL; }
        for (int64_t i = first_new - 1; i < first_new - 1 + n_new; ++i) affected[i] = 1;      // the new cells themselves
        return HVB_OK;
    }

    // interface areas aligned with the neighbour lists (hvb_geometry.cuh, first facet fixed)
    int cell_areas(double* area) override {
        if (merged || perturbed) { err = "hvb_cell_areas: not available on a context whose cloud was resolved from non-general position"; return HVB_ESTATE; }
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        if (!area) { err = "null output"; return HVB_EINVAL; }
        if (seed_prefix > 0) { err = "interface areas need all vertices of the cells: not available after a search with seed vertices"; return HVB_ESTATE; }
        int rc = build_neighbors(); if (rc) return rc;
        CK(cudaSetDevice(prm.device));
        const long long n_list = periodic ? n_user : n;
        const long long tot = std::max<long long>(nb_total, 1);
        CK(vol_acc.ensure(tot)); CK(vol_dev.ensure(tot)); CK(vol_sat.ensure(tot));
        CK(cudaMemsetAsync(vol_acc.p, 0, (size_t)tot * sizeof(long long), stream));
        CK(cudaMemsetAsync(vol_sat.p, 0, (size_t)tot, stream));
        double fact = 1.0;
        for (int k = 2; k <= D - 1; ++k) fact *= k;
        const double scale = ldexp(1.0, 52) / (pow(dv.ext, (double)(D - 1)) * fact);
        if (nvert > 0 && nb_total > 0) {
            k_cell_areas<D><<<blocks_for(nvert, 128), 128, 0, stream>>>(out_sig[res].p, (u32)nvert, xs_in.p, (long long)n, n_list, planes.p,
                                                                       nb_off.p, nb_ids.p, scale, vol_acc.p, vol_sat.p);
            ++launches;
        }
        k_volumes_finish<<<blocks_for(tot, 256), 256, 0, stream>>>(vol_acc.p, 1.0 / (scale * fact), vol_dev.p, tot, vol_sat.p); ++launches;
        if (nrays > 0 && nb_total > 0) {
            k_areas_unbounded<<<blocks_for(nrays, 256), 256, 0, stream>>>(ray_edge.p, (long long)nrays, D, n_list, nb_off.p, nb_ids.p, vol_dev.p);
            ++launches;
        }
        if (nb_total > 0) CK(cudaMemcpyAsync(area, vol_dev.p, (size_t)nb_total * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        st.kernel_launches = launches;
        return HVB_OK;
    }

    // area and first moment of every interface, aligned with the neighbour lists (hvb_cell_area_moments)
    int cell_area_moments(double* area, double* first) override {
        if (merged || perturbed) { err = "hvb_cell_area_moments: not available on a context whose cloud was resolved from non-general position"; return HVB_ESTATE; }
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        if (seed_prefix > 0) { err = "interface integrals need all vertices of the cells: not available after a search with seed vertices"; return HVB_ESTATE; }
        if (std::max(1, prm.world) > 1) { err = "hvb_cell_area_moments runs on a single-GPU context"; return HVB_EINVAL; }
        int rc = build_neighbors(); if (rc) return rc;
        CK(cudaSetDevice(prm.device));
        const long long n_list = periodic ? n_user : n;
        const long long tot = std::max<long long>(nb_total, 1);
        CK(vol_acc.ensure((size_t)tot * (1 + D))); CK(mom_dev.ensure((size_t)tot * (1 + D))); CK(vol_sat.ensure(tot));
        CK(cudaMemsetAsync(vol_acc.p, 0, (size_t)tot * (1 + D) * sizeof(long long), stream));
        CK(cudaMemsetAsync(vol_sat.p, 0, (size_t)tot, stream));
        double fact = 1.0;
        for (int k = 2; k <= D - 1; ++k) fact *= k;
        const double s0 = ldexp(1.0, 52) / (pow(dv.ext, (double)(D - 1)) * fact), s1 = s0 / dv.ext;
        if (nvert > 0 && nb_total > 0) {
            k_cell_area_moments<D><<<blocks_for(nvert, 128), 128, 0, stream>>>(out_sig[res].p, (u32)nvert, xs_in.p, (long long)n, n_list, planes.p,
                                                                              nb_off.p, nb_ids.p, s0, s1, vol_acc.p, vol_sat.p);
            ++launches;
        }
        double* darea = mom_dev.p; double* dfirst = mom_dev.p + tot;
        k_area_moments_finish<D><<<blocks_for(tot, 128), 128, 0, stream>>>(vol_acc.p, xs_in.p, nb_off.p, n_list, (long long)nb_total, 1.0 / (s0 * fact), 1.0 / (s1 * fact), darea, dfirst, vol_sat.p); ++launches;
        if (nrays > 0 && nb_total > 0) {
            k_area_moments_unbounded<<<blocks_for(nrays, 256), 256, 0, stream>>>(ray_edge.p, (long long)nrays, D, n_list, nb_off.p, nb_ids.p, darea, dfirst);
            ++launches;
        }
        if (nb_total > 0 && area) CK(cudaMemcpyAsync(area, darea, (size_t)nb_total * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (nb_total > 0 && first) CK(cudaMemcpyAsync(first, dfirst, (size_t)nb_total * D * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        st.kernel_launches = launches;
        return HVB_OK;
    }

    // expected number of periodic images of a vertex that touch caller generators (capacity estimate only)
    double periodic_versions() const {
        static const double cd[7] = {0, 0, 3.14159265358979, 4.18879020478639, 4.93480220054468, 5.26378901391432, 5.16771278004997};
        double vol = 1.0;
        for (int k = 0; k < D; ++k) vol *= std::max(dv.g[k] * dv.h[k], 1e-300);
        const double rtyp = pow(vol / (double)std::max<int64_t>(n, 1), 1.0 / D) * pow((double)D / cd[D], 1.0 / D);
        double stay = 1.0;
        for (int i = 0; i < halo.npairs; ++i) stay *= 1.0 - std::min(1.0, 2.0 * rtyp / pair_width[i]);
        return std::min((double)(D + 1), 1.0 + D * (1.0 - stay));
    }

    int search_once(const int64_t* cells, int64_t ncells_in, const int64_t* seed_sig, const double* seed_r, int64_t nseed, int stride) {
        if (nseed < 0 || (nseed > 0 && (!seed_sig || !seed_r || stride < D + 1))) { err = "bad seed vertex arguments"; return HVB_EINVAL; }
        CK(cudaSetDevice(prm.device));
        have_result = false; staged = false; staged_r = staged_sig64 = staged_sig32 = false; nb_staged32 = false; nb_total = -1; have_flags = false;
        st.exchange_bytes = 0;
        int64_t cap = prm.vertex_capacity > 0 ? prm.vertex_capacity
                      : (periodic ? (int64_t)(estimate_vertices(D, n_user, P) * periodic_versions()) : estimate_vertices(D, n, P));
        const int world = std::max(1, prm.world), rank = std::min(std::max(0, prm.rank), world - 1);
        // a slab context stores what touches its slab: 1/world of the vertices plus the layer both neighbours find too
        // (tables are memset every search: over-sizing them by `world` is what made the replicated part of a step grow)
        if (world > 1 && cells == nullptr && prm.vertex_capacity <= 0) cap = (int64_t)(cap * std::min(1.0, (D <= 3 ? 1.8 : 3.0) / world + 0.04)) + 4096;
        if (vcap >= cap) cap = vcap;
        int retries = 0;
        launches = 0;
        size_t n_ev = 0;
        int64_t rounds = 0, items = 0, expand_launches = 0;
        CK(cudaEventRecord(ev_a, stream));
        for (;;) {
            if (cap != vcap || !vsig.p) { int rc = alloc_tables(cap); if (rc) return rc; }
            CK(cudaMemsetAsync(vtab.p, 0, (dv.vmask + 1) * sizeof(u64), stream));
            CK(cudaMemsetAsync(etab.p, 0, (dv.emask + 1) * sizeof(u64), stream));
            CK(cudaMemsetAsync(has_vertex.p, 0, n, stream));
            CK(cudaMemsetAsync(ctr.p, 0, sizeof(Counters), stream));
            CK(cudaMemsetAsync(sc.p, 0, sizeof(Scalars), stream));
            if (cells == nullptr) {
                k_fill_active_owner<<<blocks_for(n, 256), 256, 0, stream>>>(active.p, owner_ptr, perm.p, (int)n, (int)(periodic ? n_user : n), rank);
                ++launches;
            } else {
                CK(cells_dev.ensure(ncells_in));
                CK(cudaMemcpyAsync(cells_dev.p, cells, ncells_in * sizeof(long long), cudaMemcpyHostToDevice, stream));
                CK(cudaMemsetAsync(active.p, 0, n, stream));
                k_mark_cells<<<blocks_for(ncells_in, 256), 256, 0, stream>>>(cells_dev.p, ncells_in, inv.p, active.p, (int)n); ++launches;
            }
            // seeds: one descent every `stride` generators of the sorted order
            int sstride = prm.seed_stride;
            if (sstride <= 0) {
                // one descent per 8 explored cells (measured on C2: 8 beats 4, 16 and 32; profiles/r1_sweeps_session2.md)
                const int64_t n_act = periodic ? n_user : n;
                int64_t want = std::min<int64_t>(std::max<int64_t>(n_act / 8, 2048), 65536);
                sstride = (int)std::max<int64_t>(1, n_act / want);
            }
            int nseeds = (int)((n + sstride - 1) / sstride);
            int cur = 0;
            if (persistent) CK(cudaMemsetAsync(q[0].p, 0xff, (size_t)qcap * sizeof(u64), stream));   // nothing published yet
            CK(cudaEventRecord(ev_s0, stream));
            seed_prefix = 0;
            if (nseed > 0) {
                // the mesh already holds vertices: they become the first records; cells they do not reach are seeded by
                // descents in the re-seed pass below (sysvoronoi.jl:394-429)
                CK(seed_sig_dev.ensure((size_t)nseed * stride)); CK(seed_r_dev.ensure((size_t)nseed * D));
                CK(cudaMemcpyAsync(seed_sig_dev.p, seed_sig, (size_t)nseed * stride * 8, cudaMemcpyHostToDevice, stream));
                CK(cudaMemcpyAsync(seed_r_dev.p, seed_r, (size_t)nseed * D * 8, cudaMemcpyHostToDevice, stream));
                k_insert_seeds<D><<<blocks_for(nseed, 128), 128, 0, stream>>>(dv, seed_sig_dev.p, seed_r_dev.p, nseed, stride, inv.p,
                                                                             q[cur].p, &sc.p->rnd[cur].qcount, qcap, &sc.p->pflags + 0);
                ++launches;
                int rcs = read_scalars(); if (rcs) return rcs;
                if (h_sc.p->pflags) { err = "seed vertices must be general vertices (dim+1 distinct, valid ids)"; return HVB_EINVAL; }
                seed_prefix = h_sc.p->vcount;
            } else
            launch_seed(nullptr, nseeds, sstride, cur);
            CK(cudaEventRecord(ev_s1, stream));
            if (debug) fprintf(stderr, "[hvb] seeds=%d stride=%d G=%d vcap=%lld ncells=%lld\n", nseeds, sstride, G, (long long)vcap, (long long)ncells);
            bool overflow = false;
            u32 last_uns = 0xffffffffu, last_vcount = 0;
            if (persistent) {
                // ---- single-launch walk: seeds were appended to q[0]; k_walk drains and extends it ----------------
                WalkQueue wq;
                wq.q = q[0].p; wq.tail = &sc.p->rnd[0].qcount; wq.head = &sc.p->q_head; wq.done = &sc.p->q_done;
                wq.abort = &sc.p->q_abort; wq.cap = qcap; wq.stop_on_degenerate = report_degenerate() ? 1u : 0u;
                for (;;) {
                    cudaEvent_t e0 = pool_event(n_ev++), e1 = pool_event(n_ev++);
                    CK(cudaEventRecord(e0, stream));
                    int rcw = launch_walk(wq); if (rcw) return rcw;
                    CK(cudaEventRecord(e1, stream));
                    ++launches; ++expand_launches; ++rounds;
                    // cells of this context without a vertex get their own descent (rarely any)
                    CK(cudaMemsetAsync(&sc.p->unseeded, 0, sizeof(u32), stream));
                    k_unseeded<D><<<blocks_for(n, 256), 256, 0, stream>>>(dv, unseeded_list.p, &sc.p->unseeded); ++launches;
                    int rc = read_scalars(); if (rc) return rc;
                    items = h_sc.p->q_done;
                    if (h_sc.p->q_abort) { err = "walk kernel timed out (internal error)"; return HVB_ECUDA; }
                    if ((h_ctr.p->degenerate > 0 || (h_ctr.p->flags & FLAG_DEGEN)) && report_degenerate()) {
                        st.degenerate = (int64_t)std::max<u64>(h_ctr.p->degenerate, 1);
                        err = "non-general position: a vertex with more than dim+1 cospherical generators was met";
                        return HVB_EDEGENERATE;
                    }
                    if (h_sc.p->pflags || (h_ctr.p->flags & FLAG_OVERFLOW_MASK)) { overflow = true; break; }
                    u32 uns = h_sc.p->unseeded;
                    if (uns == 0) break;
                    if (uns == last_uns && h_sc.p->vcount == last_vcount) break;
                    last_uns = uns; last_vcount = h_sc.p->vcount;
                    if (debug) fprintf(stderr, "[hvb] reseeding %u empty cells\n", uns);
                    // every entry below the old tail is processed: tickets restart there
                    u32 tl = h_sc.p->rnd[0].qcount;
                    CK(cudaMemcpyAsync(&sc.p->q_head, &tl, sizeof(u32), cudaMemcpyHostToDevice, stream));
                    launch_seed(unseeded_list.p, (int)uns, 1, 0);
                    if (rounds > 1000) { err = "search does not terminate"; return HVB_EINCOMPLETE; }
                }
            } else
            for (;;) {
                int rc = read_scalars(); if (rc) return rc;
                if (h_sc.p->pflags || (h_ctr.p->flags & FLAG_OVERFLOW_MASK)) { overflow = true; break; }
                if (h_ctr.p->degenerate > 0 && report_degenerate()) {
                    // non-general position (edgeiterate.jl territory): stop at once instead of walking a corrupt frontier
                    st.degenerate = (int64_t)h_ctr.p->degenerate;
                    err = "non-general position: a vertex with more than dim+1 cospherical generators was met";
                    return HVB_EDEGENERATE;
                }
                u32 cnt = h_sc.p->rnd[cur].qcount;
                if (cnt > 0) {
                    int nxt = 1 - cur;
                    CK(cudaMemsetAsync(&sc.p->rnd[nxt], 0, sizeof(Round), stream));
                    cudaEvent_t e0 = pool_event(n_ev++), e1 = pool_event(n_ev++);
                    CK(cudaEventRecord(e0, stream));
                    if (debug) fprintf(stderr, "[hvb] round %lld frontier=%u vertices=%u\n", (long long)rounds, cnt, h_sc.p->vcount);
                    launch_expand(cnt, cur, nxt);
                    CK(cudaEventRecord(e1, stream));
                    ++launches; ++expand_launches; ++rounds; items += cnt;
                    cur = nxt;
                    continue;
                }
                // frontier drained: cells of this context without a vertex get their own descent
                CK(cudaMemsetAsync(&sc.p->unseeded, 0, sizeof(u32), stream));
                k_unseeded<D><<<blocks_for(n, 256), 256, 0, stream>>>(dv, unseeded_list.p, &sc.p->unseeded); ++launches;
                rc = read_scalars(); if (rc) return rc;
                u32 uns = h_sc.p->unseeded;
                if (uns == 0) break;
                if (uns == last_uns && h_sc.p->vcount == last_vcount) break;   // descents keep failing: give up (HVB_EINCOMPLETE)
                last_uns = uns; last_vcount = h_sc.p->vcount;
                CK(cudaMemsetAsync(&sc.p->rnd[cur], 0, sizeof(Round), stream));
                if (debug) fprintf(stderr, "[hvb] reseeding %u empty cells\n", uns);
                launch_seed(unseeded_list.p, (int)uns, 1, cur);
                ++rounds;
                if (rounds > 100000) { err = "search does not terminate"; return HVB_EINCOMPLETE; }
            }
            if (!overflow) break;
            if (++retries > 6) { err = "capacity exhausted after 6 retries"; return HVB_ENOMEM; }
            cap = vcap * 2;
        }
        const bool by_slab = world > 1 && cells == nullptr;
        own_ptr = nullptr;
        if (by_slab) {
            CK(own_mask.ensure(n));
            k_own_mask<<<blocks_for(n, 256), 256, 0, stream>>>(perm.p, (int)n, owner_ptr, rank, own_mask.p); ++launches;
            own_ptr = own_mask.p;
        }
        CK(cudaEventRecord(ev_b, stream));
        // The neighbour lists are built from the vertex records of the walk (a set of pairs needs neither the result rows
        // nor their order) on their own stream, next to k_final_rows and the radix sort of the rows.  With seed vertices
        // they must be built now: the caller's own vertices are part of the lists but not of the returned rows
        const bool want_nb = prm.neighbors || seed_prefix > 0;
        int rc = HVB_OK;
        CK(cudaStreamWaitEvent(nstream, ev_b, 0));
        CK(cudaEventRecord(ev_n0, nstream));
        if (want_nb) { rc = nb_prepare(true, nullptr); if (rc) return rc; rc = nb_enqueue(nstream); if (rc) return rc; }
        rc = finalize(by_slab); if (rc) return rc;             // rows + sort: enqueued, not waited for
        CK(cudaEventRecord(ev_c, stream));
        // the lists are finished (one host wait on THEIR stream for the list sizes, then fill + sort) while the row sort runs
        if (want_nb) { rc = nb_finish(nstream); if (rc) return rc; }
        rc = finalize_collect(by_slab); if (rc) return rc;
        // a slab result is an intermediate: it is exported to the exchange step, not staged for the host
        if (world == 1) { rc = stage(); if (rc) return rc; }
        have_result = true;
        if (prm.neighbors) { rc = stage_neighbors(); if (rc) return rc; }
        CK(cudaEventRecord(ev_n1, nstream));
        CK(cudaStreamWaitEvent(stream, ev_n1, 0));
        // ev_d: the result (rows, neighbour lists) is complete in HBM.  The page-locked staging copies run on their own
        // stream and are waited for here, outside ms_finalize: they belong to the end-to-end time, not to the search
        CK(cudaEventRecord(ev_d, stream));
        CK(cudaStreamWaitEvent(stream, ev_stage_done(), 0));
        if (!ev_sd2) CK(cudaEventCreateWithFlags(&ev_sd2, cudaEventDisableTiming));
        CK(cudaEventRecord(ev_sd2, sstream2));
        CK(cudaStreamWaitEvent(stream, ev_sd2, 0));
        if (counts_cached) CK(cudaStreamWaitEvent(stream, ev_x1, 0));
        CK(cudaEventRecord(ev_p1, stream));
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_d, ev_p1); st.ms_stage_wait = ms;
        cudaEventElapsedTime(&ms, ev_a, ev_b); st.ms_search = ms;
        cudaEventElapsedTime(&ms, ev_b, ev_d); st.ms_finalize = ms;
        cudaEventElapsedTime(&ms, ev_s0, ev_s1); st.ms_seed = ms;
        cudaEventElapsedTime(&ms, ev_n0, ev_n1); st.ms_neighbors = ms;
        cudaEventElapsedTime(&ms, ev_b, ev_c); st.ms_rows_sort = ms;
        double kms = 0;
        for (size_t i = 0; i + 1 < n_ev; i += 2) { cudaEventElapsedTime(&ms, ev_pool[i], ev_pool[i + 1]); kms += ms; }
        st.ms_expand_kernel = kms; st.expand_launches = expand_launches; st.expand_items = items;
        const Counters& c = *h_ctr.p;
        st.vertices = nvert; st.unique_vertices = nvert; st.periodic_retries = 0; st.rays = nrays; st.raycasts = (int64_t)c.raycasts; st.duplicate_hits = (int64_t)c.dup_hits;
        st.closed_skips = (int64_t)c.closed_skips; st.candidates_fp32 = (int64_t)c.cand32; st.candidates_fp64 = (int64_t)c.cand64;
        st.rows_scanned = (int64_t)c.rows; st.probe_stages = (int64_t)c.stages; st.rounds = rounds; st.seeds = (int64_t)c.seeds;
        st.degenerate = (int64_t)c.degenerate; st.kernel_launches = launches; st.capacity_retries = retries;
        have_result = true;
        if (c.degenerate > 0 && report_degenerate()) {
            err = "non-general position: a vertex with more than dim+1 cospherical generators was met"; return HVB_EDEGENERATE;
        }
        if (c.seed_fail > 0 && h_sc.p->unseeded > 0) { err = "descent failed for some cells"; return HVB_EINCOMPLETE; }
        return HVB_OK;
    }

    int id_bits() const { int b = 1; while ((1LL << b) < n + P + 1) ++b; return b; }

    // sorts `count` rows held in out_sig[0]/out_r[0] (192-bit keys in key_top/key_hi/key_lo) into out_sig[1]/out_r[1]:
    // LSD radix sort, one stable pass per 64-bit key word that carries bits
    int sort_rows(u32 count, int bits) {
        res = 0;
        if (!prm.sort_output || count == 0) return HVB_OK;
        const int total_bits = (D + 1) * bits;
        if (total_bits > 192) return HVB_OK;               // cannot happen for n < 2^27 (d = 6) / 2^31 (d <= 5)
        CK(idx[0].ensure(count)); CK(idx[1].ensure(count)); CK(key_tmp.ensure(count));
        k_iota<<<blocks_for(count, 256), 256, 0, stream>>>(idx[0].p, count); ++launches;
        const u64* words[3] = {key_lo.p, key_hi.p, key_top.p};
        int cur = 0;
        for (int w = 0; w < 3; ++w) {
            const int wbits = std::min(64, total_bits - 64 * w);
            if (wbits <= 0) break;
            const u64* keys = words[w];
            if (w > 0) {        // bring the word into the current order (key_lo is free after the first pass)
                k_gather_u64<<<blocks_for(count, 256), 256, 0, stream>>>(words[w], idx[cur].p, key_lo.p, count); ++launches;
                keys = key_lo.p;
            }
            size_t tmp_bytes = 0;
            CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, key_tmp.p, idx[cur].p, idx[1 - cur].p, (int)count, 0, wbits, stream));
            CK(cub_tmp.ensure(tmp_bytes));
            CK(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, keys, key_tmp.p, idx[cur].p, idx[1 - cur].p, (int)count, 0, wbits, stream));
            cur = 1 - cur;
        }
        k_gather_rows<D><<<blocks_for(count, 256), 256, 0, stream>>>(out_sig[0].p, out_r[0].p, idx[cur].p, out_sig[1].p, out_r[1].p, count); ++launches;
        res = 1;
        return HVB_OK;
    }

    int finalize(bool by_slab) {
        u32 nrec = std::min<u32>(h_sc.p->vcount, (u32)vcap);
        nrays = std::min<u32>(h_sc.p->ray_count, ray_cap);
        for (int i = 0; i < 2; ++i) { CK(out_sig[i].ensure((size_t)std::max<u32>(nrec, 1) * (D + 1))); CK(out_r[i].ensure((size_t)std::max<u32>(nrec, 1) * D)); }
        CK(key_top.ensure(std::max<u32>(nrec, 1))); CK(key_hi.ensure(std::max<u32>(nrec, 1))); CK(key_lo.ensure(std::max<u32>(nrec, 1)));
        int bits = id_bits();
        if (use_buckets()) {
            CK(bkt_count.ensure((size_t)n + 2)); CK(bkt_start.ensure((size_t)n + 2));
            CK(cudaMemsetAsync(bkt_count.p, 0, ((size_t)n + 2) * sizeof(int), stream));
        } else
        if (nrec > 0 && prm.sort_output) {
            CK(cudaMemsetAsync(key_top.p, 0xff, (size_t)nrec * sizeof(u64), stream));
            CK(cudaMemsetAsync(key_hi.p, 0xff, (size_t)nrec * sizeof(u64), stream));
            CK(cudaMemsetAsync(key_lo.p, 0xff, (size_t)nrec * sizeof(u64), stream));
        }
        if (nrec > 0) {
            // multi-GPU: only the vertices this rank owns; seed vertices (the caller's own) are not returned
            const int world = std::max(1, prm.world), rank = std::min(std::max(0, prm.rank), world - 1);
            k_final_rows<D><<<blocks_for(nrec, 128), 128, 0, stream>>>(dv, perm.p, nrec, bits, out_sig[0].p, out_r[0].p, key_top.p, key_hi.p, key_lo.p,
                                                                     &sc.p->out_count, &sc.p->max_var, by_slab ? owner_ptr : nullptr, rank, seed_prefix,
                                                                     prm.variance_tol, prm.break_tol, sc.p->tol_counts, (int)n_user,
                                                                     perturbed ? HVB_FLAT_TOL : 0.0, &sc.p->pad3, use_buckets() ? bkt_count.p : nullptr);
            ++launches;
        }
        if (nrays > 0) {
            CK(ray_edge.ensure((size_t)nrays * D)); CK(ray_base.ensure((size_t)nrays * D)); CK(ray_dir.ensure((size_t)nrays * D)); CK(ray_node.ensure(nrays));
            const int world = std::max(1, prm.world), rank = std::min(std::max(0, prm.rank), world - 1);
            k_final_rays<D><<<blocks_for(nrays, 128), 128, 0, stream>>>(dv, perm.p, (u32)nrays, ray_edge.p, ray_base.p, ray_dir.p, ray_node.p,
                                                                     by_slab ? owner_ptr : nullptr, rank, &sc.p->ray_out);
            ++launches;
        }
        // multi-GPU with a communicator: the shard sizes of all ranks travel now -- one ncclAllGather of one word per rank,
        // straight from the device counter, on its own stream: a collective makes the
        // fast ranks wait for the slowest one, and that wait must overlap the row sort and the neighbour lists instead of
        // standing in front of them.  The numbers are in host memory when this search returns.
        counts_cached = false;
        if (by_slab && comm && !getenv("HVB_NO_COUNT_XCHG")) {
            const int world = std::max(1, prm.world);
            if (!xstream) { CK(cudaStreamCreateWithFlags(&xstream, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&ev_x0, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ev_x1, cudaEventDisableTiming)); }
            CK(xc_counts32.ensure(world)); CK(h_xc_counts32.ensure(world));
            CK(cudaEventRecord(ev_x0, stream));
            CK(cudaStreamWaitEvent(xstream, ev_x0, 0));
            NK(Nccl::get().AllGather(&sc.p->out_count, xc_counts32.p, 1, ncclUint32, comm, xstream));
            CK(cudaMemcpyAsync(h_xc_counts32.p, xc_counts32.p, world * sizeof(u32), cudaMemcpyDeviceToHost, xstream));
            CK(cudaEventRecord(ev_x1, xstream));
            counts_cached = true;
        }
        // The sort is enqueued over ALL nrec records right behind k_final_rows -- rows it skipped (dead records, vertices
        // of other ranks, rejected ones) keep the all-ones key the arrays were filled with and end up behind the result --
        // so that no host round trip stands between the two: the row count is read while the sort runs.
        return use_buckets() ? sort_rows_bucket(nrec) : sort_rows(nrec, bits);
    }
    // d = 2: counting sort on the first generator + insertion sort inside the buckets (k_bucket_scatter / k_bucket_sort):
    // 6 launches instead of ~20 (C3, 2e6 rows: rows + sort 0.77 -> 0.55 ms).  Measured for d = 3 as well (C2): the same 0.40 ms
    // as the radix passes, and the finalize is bound by the neighbour lists on the other stream there -- not used')
    DBuf<int> bkt_count, bkt_start;
    bool use_buckets() const { return prm.sort_output && D == 2 && D * id_bits() <= 64 && !getenv("HVB_RADIX_SORT"); }
    int sort_rows_bucket(u32 nrec) {
        res = 0;
        if (nrec == 0) return HVB_OK;
        CK(idx[0].ensure(nrec));
        size_t tmp_bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, bkt_count.p, bkt_start.p, (int)(n + 1), stream));
        CK(cub_tmp.ensure(tmp_bytes));
        CK(cub::DeviceScan::ExclusiveSum(cub_tmp.p, tmp_bytes, bkt_count.p, bkt_start.p, (int)(n + 1), stream));
        CK(cudaMemsetAsync(bkt_count.p, 0, ((size_t)n + 1) * sizeof(int), stream));        // reused as the cursors
        k_bucket_scatter<<<blocks_for(nrec, 256), 256, 0, stream>>>(out_sig[0].p, D + 1, &sc.p->out_count, bkt_start.p, bkt_count.p, idx[0].p); ++launches;
        k_bucket_sort<<<blocks_for(n, 128), 128, 0, stream>>>(key_lo.p, bkt_start.p, (int)n, idx[0].p); ++launches;
        k_gather_rows<D><<<blocks_for(nrec, 256), 256, 0, stream>>>(out_sig[0].p, out_r[0].p, idx[0].p, out_sig[1].p, out_r[1].p, nrec, &sc.p->out_count); ++launches;
        res = 1;
        return HVB_OK;
    }
    // second half of finalize: waits for the row count (the neighbour lists were completed on their stream meanwhile)
    int finalize_collect(bool by_slab) {
        int rc = read_scalars(); if (rc) return rc;
        nvert = h_sc.p->out_count;
        if (by_slab) nrays = h_sc.p->ray_out;
        st.rejected = h_sc.p->tol_counts[0]; st.suboptimal = h_sc.p->tol_counts[1];
        return HVB_OK;
    }

    // device -> page-locked host staging (asynchronous; the fetch calls wait for it).  Three pieces, staged on demand:
    // coordinates, signatures as int64 (the reference's Int64 ids) and signatures as int32 (the compact wire format:
    // prm.wire32 stages this one inside hvb_search instead of the int64 form, 4 (dim+1) bytes per row less over PCIe)
    bool staged_r = false, staged_sig64 = false, staged_sig32 = false;
    DBuf<int> sig32_dev, ids32_dev;
    HBuf<int> h_sig32, h_ids32;
    bool nb_staged32 = false;
    int stage_rows(bool want64, bool want32) {
        const size_t rows = (size_t)std::max<int64_t>(nvert, 1);
        const bool do_r = !staged_r, do64 = want64 && !staged_sig64, do32 = want32 && !staged_sig32;
        if (!(do_r || do64 || do32)) return HVB_OK;
        if (do_r) CK(h_r.ensure(rows * D));
        if (do64) CK(h_sig.ensure(rows * (D + 1)));
        if (do32) { CK(h_sig32.ensure(rows * (D + 1))); CK(sig32_dev.ensure(rows * (D + 1))); }
        if (nvert > 0) {
            if (do32) { const size_t cnt = (size_t)nvert * (D + 1); k_narrow_i64<<<blocks_for((int64_t)cnt, 256), 256, 0, stream>>>(out_sig[res].p, sig32_dev.p, cnt); ++launches; }
            // on the staging stream, so that the copy overlaps whatever the compute stream does next (neighbour lists)
            CK(cudaEventRecord(ev_stage, stream));
            CK(cudaStreamWaitEvent(sstream, ev_stage, 0));
            if (do32) CK(cudaMemcpyAsync(h_sig32.p, sig32_dev.p, (size_t)nvert * (D + 1) * sizeof(int), cudaMemcpyDeviceToHost, sstream));
            if (do64) CK(cudaMemcpyAsync(h_sig.p, out_sig[res].p, (size_t)nvert * (D + 1) * sizeof(long long), cudaMemcpyDeviceToHost, sstream));
            if (do_r) CK(cudaMemcpyAsync(h_r.p, out_r[res].p, (size_t)nvert * D * sizeof(double), cudaMemcpyDeviceToHost, sstream));
        }
        staged_r = true; staged_sig64 |= want64; staged_sig32 |= want32;
        staged = staged_sig64;
        return HVB_OK;
    }
    int stage() { return stage_rows(!prm.wire32, prm.wire32 != 0); }

    cudaEvent_t ev_sd = nullptr, ev_sd2 = nullptr;
    cudaEvent_t ev_stage_done() {          // an event on the staging stream that marks "everything staged so far is in host memory"
        if (!ev_sd) cudaEventCreateWithFlags(&ev_sd, cudaEventDisableTiming);
        cudaEventRecord(ev_sd, sstream);
        return ev_sd;
    }
    int counts(int64_t* nv, int64_t* nr, int64_t* msl) override {
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        if (nv) *nv = merged ? m_nvert : nvert;
        if (nr) *nr = nrays;
        if (msl) *msl = merged ? m_maxlen : D + 1;
        return HVB_OK;
    }
    // fixed-width calls on a merged result: fine while every vertex still has dim + 1 generators
    int merged_fixed(const char* what) {
        if (merged && m_maxlen > D + 1) {
            err = std::string(what) + ": the mesh holds vertices with more than dim+1 generators (non-general position): use hvb_fetch_vertices_var";
            return HVB_ESTATE;
        }
        return HVB_OK;
    }
    int view_vertices(const int64_t** sig, const double** r, int64_t* nv) override {
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        { int rcm = merged_fixed("hvb_view_vertices"); if (rcm) return rcm; }
        CK(cudaSetDevice(prm.device));
        { int rc = stage_rows(true, false); if (rc) return rc; }
        CK(cudaStreamSynchronize(stream));
        CK(cudaStreamSynchronize(sstream));
        *sig = (const int64_t*)h_sig.p; *r = h_r.p; *nv = nvert;
        return HVB_OK;
    }
    int view_vertices32(const int32_t** sig, const double** r, int64_t* nv) override {
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        { int rcm = merged_fixed("hvb_view_vertices32"); if (rcm) return rcm; }
        if (n + P >= 0x7fffffffLL) { err = "ids do not fit 32 bits"; return HVB_EINVAL; }
        CK(cudaSetDevice(prm.device));
        { int rc = stage_rows(false, true); if (rc) return rc; }
        CK(cudaStreamSynchronize(stream));
        CK(cudaStreamSynchronize(sstream));
        *sig = (const int32_t*)h_sig32.p; *r = h_r.p; *nv = nvert;
        return HVB_OK;
    }
    int fetch_vertices(int64_t* sig, double* r) override {
        if (merged) {
            int rcm = merged_fixed("hvb_fetch_vertices"); if (rcm) return rcm;
            if (sig && !m_ids.empty()) memcpy(sig, m_ids.data(), m_ids.size() * sizeof(int64_t));
            if (r && !m_r.empty()) memcpy(r, m_r.data(), m_r.size() * sizeof(double));
            return HVB_OK;
        }
        const int64_t* s; const double* rr; int64_t nv;
        int rc = view_vertices(&s, &rr, &nv); if (rc) return rc;
        if (sig) memcpy(sig, s, (size_t)nv * (D + 1) * sizeof(int64_t));
        if (r) memcpy(r, rr, (size_t)nv * D * sizeof(double));
        return HVB_OK;
    }
    // rows [first, first + count) straight from the device (no staging of the whole result): the shard of a rank
    int fetch_vertices_range(int64_t first, int64_t count, int64_t* sig, double* r) override {
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        { int rcm = merged_fixed("hvb_fetch_vertices_range"); if (rcm) return rcm; }
        if (first < 0 || count < 0 || first + count > nvert) { err = "row range out of bounds"; return HVB_EINVAL; }
        CK(cudaSetDevice(prm.device));
        if (count > 0) {
            if (sig) CK(cudaMemcpyAsync(sig, out_sig[res].p + (size_t)first * (D + 1), (size_t)count * (D + 1) * 8, cudaMemcpyDeviceToHost, stream));
            if (r) CK(cudaMemcpyAsync(r, out_r[res].p + (size_t)first * D, (size_t)count * D * 8, cudaMemcpyDeviceToHost, stream));
        }
        CK(cudaStreamSynchronize(stream));
        return HVB_OK;
    }
    int fetch_rays(int64_t* edge, double* base, double* dir, int64_t* node) override {
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        CK(cudaSetDevice(prm.device));
        if (nrays == 0) return HVB_OK;
        if (edge) CK(cudaMemcpyAsync(edge, ray_edge.p, (size_t)nrays * D * 8, cudaMemcpyDeviceToHost, stream));
        if (base) CK(cudaMemcpyAsync(base, ray_base.p, (size_t)nrays * D * 8, cudaMemcpyDeviceToHost, stream));
        if (dir) CK(cudaMemcpyAsync(dir, ray_dir.p, (size_t)nrays * D * 8, cudaMemcpyDeviceToHost, stream));
        if (node) CK(cudaMemcpyAsync(node, ray_node.p, (size_t)nrays * 8, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return HVB_OK;
    }

    // ---- neighbour lists (neighbors_of_cell_new, neighbors.jl:219-262) ------------------------------------------
    // Two sources: the vertex records of the walk (raw: inside hvb_search, next to k_final_rows and the row sort) or
    // the result rows (a request after the search, or after a multi-GPU merge installed other rows).
    u64 nb_want = 0;
    bool nb_raw = false;
    const long long* nb_rows = nullptr;
    int nb_prepare(bool raw, const long long* rows) {
        nb_staged = false; nb_staged32 = false; nb_off_staged = false; nb_raw = raw; nb_rows = rows;
        CK(deg.ensure(n)); CK(ncur.ensure(n)); CK(nb_off.ensure(n + 1));
        static const double nb_est[7] = {0, 0, 8, 20, 48, 120, 320};
        // periodic contexts build the lists of the caller's cells only (n_user == n otherwise)
        const long long n_list = periodic ? n_user : n;
        const double nrows = raw ? (double)std::min<u32>(h_sc.p->vcount, (u32)vcap) : (double)nvert;
        // unordered pairs: a list entry of an interior cell is stored once for two cells, so about n * nb_est / 2 pairs;
        // slots = 2 x that estimate (the estimate itself is ~1.3 x the Poisson-Voronoi mean): load <= 0.4, and the
        // memset + the fill pass touch a quarter of what an entry-per-list-element table would need
        // (slab contexts build the lists of their own cells only: 1/world of the cells plus the pairs that cross the slab faces)
        const double own_frac = own_ptr ? std::min(1.0, 2.2 / std::max(1, prm.world) + 0.05) : 1.0;
        nb_want = next_pow2((u64)std::min(nrows * D * (D + 1) / 2.0 * 2.0, (double)n_list * own_frac * nb_est[D]) + 1024);
        size_t tmp_bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, nb_off.p, nb_off.p, (int)(n + 1), stream));
        CK(nb_cub_tmp.ensure(tmp_bytes));
        nb_tmp_bytes = tmp_bytes;
        return HVB_OK;
    }
    size_t nb_tmp_bytes = 0;
    // queues pair set + degrees + offsets + the publication of {overflow flag, total} on `ns`; no host wait
    int nb_enqueue(cudaStream_t ns) {
        const long long n_list = periodic ? n_user : n;
        CK(ptab.ensure(nb_want));
        CK(cudaMemsetAsync(ptab.p, 0, nb_want * sizeof(u64), ns));
        CK(cudaMemsetAsync(deg.p, 0, n * sizeof(u32), ns));
        CK(cudaMemsetAsync(nbsc.p, 0, sizeof(NbScalars), ns));
        if (nb_raw) {
            const u32 nrec = std::min<u32>(h_sc.p->vcount, (u32)vcap);
            if (nrec > 0) { k_pairs_raw<D><<<blocks_for(nrec, 128), 128, 0, ns>>>(dv, perm.p, nrec, n_list, ptab.p, nb_want - 1, deg.p, &nbsc.p->pflags, own_ptr); ++launches; }
        } else if (nvert > 0) {
            k_pairs<D><<<blocks_for(nvert, 128), 128, 0, ns>>>(nb_rows, (u32)nvert, n_list, ptab.p, nb_want - 1, deg.p, &nbsc.p->pflags, own_ptr); ++launches;
        }
        // offsets = exclusive scan of the degrees (as int64); one host round trip brings the overflow flag and the total
        k_u32_to_i64<<<blocks_for(n, 256), 256, 0, ns>>>(deg.p, nb_off.p, n); ++launches;
        CK(cudaMemsetAsync(nb_off.p + n, 0, sizeof(long long), ns));
        size_t tmp_bytes = nb_tmp_bytes;
        CK(cub::DeviceScan::ExclusiveSum(nb_cub_tmp.p, tmp_bytes, nb_off.p, nb_off.p, (int)(n + 1), ns));
        k_publish<<<1, 32, 0, ns>>>((const u32*)nbsc.p, (u32*)h_nbsc.p, (int)(sizeof(NbScalars) / 4), nullptr, nullptr, 0, nb_off.p + n, h_nbtotal.p);
        ++launches;
        return HVB_OK;
    }
    // waits for nb_enqueue, repeats it with a larger table if the pair set overflowed, then fills and sorts the lists
    int nb_finish(cudaStream_t ns) {
        const long long n_list = periodic ? n_user : n;
        long long total = 0;
        for (int attempt = 0;; ++attempt) {
            CK(cudaStreamSynchronize(ns));
            total = *h_nbtotal.p;
            if (!(h_nbsc.p->pflags & 8u)) break;
            if (attempt == 7) { err = "neighbour pair table overflow"; return HVB_ENOMEM; }
            nb_want *= 4;
            int rc = nb_enqueue(ns); if (rc) return rc;
        }
        CK(nb_ids.ensure(std::max<long long>(total, 1)));
        CK(cudaMemsetAsync(ncur.p, 0, n * sizeof(u32), ns));
        k_pair_fill<<<blocks_for((int64_t)nb_want, 256), 256, 0, ns>>>(ptab.p, nb_want, n_list, nb_off.p, ncur.p, nb_ids.p, own_ptr); ++launches;
        k_sort_lists<<<blocks_for(n, 128), 128, 0, ns>>>(nb_off.p, nb_ids.p, n); ++launches;
        CK(cudaGetLastError());          // no host wait here: the staging copy / the fetch calls order themselves behind the stream
        CK(cudaEventRecord(ev_nb, ns));
        CK(cudaStreamWaitEvent(stream, ev_nb, 0));       // whatever the compute stream does next sees the lists
        nb_total = total;
        st.kernel_launches = launches;
        return HVB_OK;
    }
    // lists from the current result rows, on the compute stream (requests after the search)
    int build_neighbors() {
        if (nb_total >= 0) return HVB_OK;
        CK(cudaSetDevice(prm.device));
        // a slab result holds only the rows this rank OWNS; the lists of its own cells need every vertex it FOUND, so
        // they come from the vertex records of the walk (still intact: the next search / set_points drops the result)
        int rc = nb_prepare(own_ptr != nullptr, out_sig[res].p); if (rc) return rc;
        rc = nb_enqueue(stream); if (rc) return rc;
        return nb_finish(stream);
    }
    int neighbor_count(int64_t* total) override {
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        if (merged) { int rcm = build_merged_neighbors(); if (rcm) return rcm; *total = (int64_t)m_nb_ids.size(); return HVB_OK; }
        int rc = build_neighbors(); if (rc) return rc;
        *total = nb_total;
        return HVB_OK;
    }
    bool nb_off_staged = false;
    int stage_neighbors() { return stage_neighbors_as(!prm.wire32, prm.wire32 != 0); }
    int stage_neighbors_as(bool want64, bool want32) {
        const bool do64 = want64 && !nb_staged, do32 = want32 && !nb_staged32;
        if (!(do64 || do32)) return HVB_OK;
        CK(h_nb_off.ensure(n + 1));
        if (do64) CK(h_nb_ids.ensure(std::max<int64_t>(nb_total, 1)));
        if (do32) {
            CK(h_ids32.ensure(std::max<int64_t>(nb_total, 1))); CK(ids32_dev.ensure(std::max<int64_t>(nb_total, 1)));
            // the narrowing runs behind the list build, on the stream that built the lists last
            if (nb_total > 0) {
                CK(cudaStreamWaitEvent(sstream2, ev_nb, 0));
                k_narrow_i64<<<blocks_for(nb_total, 256), 256, 0, sstream2>>>(nb_ids.p, ids32_dev.p, (size_t)nb_total); ++launches;
            }
        }
        CK(cudaStreamWaitEvent(sstream2, ev_nb, 0));
        if (!nb_off_staged) CK(cudaMemcpyAsync(h_nb_off.p, nb_off.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, sstream2));
        if (do64 && nb_total > 0) CK(cudaMemcpyAsync(h_nb_ids.p, nb_ids.p, (size_t)nb_total * 8, cudaMemcpyDeviceToHost, sstream2));
        if (do32 && nb_total > 0) CK(cudaMemcpyAsync(h_ids32.p, ids32_dev.p, (size_t)nb_total * 4, cudaMemcpyDeviceToHost, sstream2));
        nb_off_staged = true; nb_staged |= want64; nb_staged32 |= want32;
        return HVB_OK;
    }
    int view_neighbors32(const int64_t** off, const int32_t** ids, int64_t* total) override {
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        if (merged) { err = "hvb_view_neighbors32: not available for a mesh resolved from non-general position (use hvb_view_neighbors)"; return HVB_ESTATE; }
        if (n + P >= 0x7fffffffLL) { err = "ids do not fit 32 bits"; return HVB_EINVAL; }
        CK(cudaSetDevice(prm.device));
        int rc = build_neighbors(); if (rc) return rc;
        rc = stage_neighbors_as(false, true); if (rc) return rc;
        CK(cudaStreamSynchronize(stream));
        CK(cudaStreamSynchronize(sstream2));
        *off = (const int64_t*)h_nb_off.p; *ids = (const int32_t*)h_ids32.p; *total = nb_total;
        return HVB_OK;
    }
    int view_neighbors(const int64_t** off, const int64_t** ids, int64_t* total) override {
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        if (merged) {
            int rcm = build_merged_neighbors(); if (rcm) return rcm;
            *off = m_nb_off.data(); *ids = m_nb_ids.data(); *total = (int64_t)m_nb_ids.size();
            return HVB_OK;
        }
        CK(cudaSetDevice(prm.device));
        int rc = build_neighbors(); if (rc) return rc;
        rc = stage_neighbors_as(true, false); if (rc) return rc;
        CK(cudaStreamSynchronize(stream));
        CK(cudaStreamSynchronize(sstream2));
        *off = (const int64_t*)h_nb_off.p; *ids = (const int64_t*)h_nb_ids.p; *total = nb_total;
        return HVB_OK;
    }
    int fetch_neighbors(int64_t* off, int64_t* ids) override {
        const int64_t *o, *i; int64_t tot;
        int rc = view_neighbors(&o, &i, &tot); if (rc) return rc;
        if (off) memcpy(off, o, (size_t)(n + 1) * 8);
        if (ids && tot > 0) memcpy(ids, i, (size_t)tot * 8);
        return HVB_OK;
    }

    // ---- in-library collective: one NCCL communicator per context (process-per-GPU callers: hvb_comm_init; the
    // single-process multi-GPU context of hvb_create_multi attaches communicators made by ncclCommInitAll) -------------
    ncclComm_t comm = nullptr;
    bool comm_owned = false;
    cudaStream_t xstream = nullptr;
    cudaEvent_t ev_x0 = nullptr, ev_x1 = nullptr;
    DBuf<long long> xc_counts;
    HBuf<long long> h_xc_counts;
    DBuf<u32> xc_counts32;
    HBuf<u32> h_xc_counts32;
    bool counts_cached = false;       // h_xc_counts32 holds the shard sizes of the current result (filled inside hvb_search)
    DBuf<int> xs_sig32, xr_sig32;
    DBuf<double> xs_r, xr_r, xc_red;
    HBuf<double> h_xc_red;
    int need_nccl() {
        Nccl& N = Nccl::get();
        if (!N.ok()) { err = "NCCL is not available: " + N.error; return HVB_ENCCL; }
        return HVB_OK;
    }
    int comm_init(const void* id128) override {
        if (!id128) { err = "null NCCL id"; return HVB_EINVAL; }
        int rc = need_nccl(); if (rc) return rc;
        CK(cudaSetDevice(prm.device));
        const int world = std::max(1, prm.world), rank = std::min(std::max(0, prm.rank), world - 1);
        if (comm && comm_owned) Nccl::get().CommDestroy(comm);
        comm = nullptr;
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        NK(Nccl::get().CommInitRank(&comm, world, id, rank));
        comm_owned = true;
        return HVB_OK;
    }
    int comm_attach(void* c) override { comm = (ncclComm_t)c; comm_owned = false; return HVB_OK; }
    // max over the ranks of a few doubles (periodic contexts agree on the next halo margin with this)
    int allreduce_max(double* v, int cnt) {
        const int world = std::max(1, prm.world);
        if (world == 1) return HVB_OK;
        if (!comm) { err = "a periodic search on several GPUs needs a communicator (hvb_comm_init / hvb_create_multi)"; return HVB_ENCCL; }
        CK(xc_red.ensure(cnt)); CK(h_xc_red.ensure(cnt));
        memcpy(h_xc_red.p, v, cnt * sizeof(double));
        CK(cudaMemcpyAsync(xc_red.p, h_xc_red.p, cnt * sizeof(double), cudaMemcpyHostToDevice, stream));
        NK(Nccl::get().AllReduce(xc_red.p, xc_red.p, cnt, ncclFloat64, ncclMax, comm, stream));
        CK(cudaMemcpyAsync(h_xc_red.p, xc_red.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        memcpy(v, h_xc_red.p, cnt * sizeof(double));
        return HVB_OK;
    }
    // counts[k] = rows rank k owns (all-gather of one word per rank; one host wait)
    int exchange_counts(int64_t* counts) override {
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        const int world = std::max(1, prm.world);
        if (world == 1) { if (counts) counts[0] = nvert; return HVB_OK; }
        if (!comm) { err = "no communicator: call hvb_comm_init first"; return HVB_ENCCL; }
        CK(cudaSetDevice(prm.device));
        CK(xc_counts.ensure(world + 1)); CK(h_xc_counts.ensure(world + 1));
        if (counts_cached) {
            // gathered inside hvb_search, next to the row sort; the device copy the unpack kernel reads is refreshed
            for (int k = 0; k < world; ++k) h_xc_counts.p[k] = (long long)h_xc_counts32.p[k];
            CK(cudaMemcpyAsync(xc_counts.p, h_xc_counts.p, world * sizeof(long long), cudaMemcpyHostToDevice, stream));
            if (counts) for (int k = 0; k < world; ++k) counts[k] = h_xc_counts.p[k];
            return HVB_OK;
        }
        h_xc_counts.p[world] = nvert;
        CK(cudaMemcpyAsync(xc_counts.p + world, h_xc_counts.p + world, sizeof(long long), cudaMemcpyHostToDevice, stream));
        NK(Nccl::get().AllGather(xc_counts.p + world, xc_counts.p, 1, ncclInt64, comm, stream));
        CK(cudaMemcpyAsync(h_xc_counts.p, xc_counts.p, world * sizeof(long long), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (counts) for (int k = 0; k < world; ++k) counts[k] = h_xc_counts.p[k];
        return HVB_OK;
    }
    // Replaces this rank's owned rows by the rows of ALL ranks (parallelmesh.jl's shared store): the owned sets are
    // disjoint and sorted, so the concatenation in rank order is the deduplicated global vertex list.  Wire format:
    // (D + 1) int32 ids + D doubles per row, padded to the largest count, one fused NCCL group of two all-gathers.
    int allgather() override {
        const int world = std::max(1, prm.world);
        if (world == 1) return have_result ? HVB_OK : (err = "no search result", HVB_ESTATE);
        int rc = exchange_counts(nullptr); if (rc) return rc;
        long long cap = 1, total = 0;
        for (int k = 0; k < world; ++k) { cap = std::max(cap, h_xc_counts.p[k]); total += h_xc_counts.p[k]; }
        if (total >= (1LL << 31)) { err = "merged vertex list beyond 2^31 rows"; return HVB_ENOMEM; }
        CK(xs_sig32.ensure((size_t)cap * (D + 1))); CK(xs_r.ensure((size_t)cap * D));
        CK(xr_sig32.ensure((size_t)world * cap * (D + 1))); CK(xr_r.ensure((size_t)world * cap * D));
        if (nvert > 0) {
            const size_t cnt = (size_t)nvert * (D + 1);
            k_narrow_i64<<<blocks_for((int64_t)cnt, 256), 256, 0, stream>>>(out_sig[res].p, xs_sig32.p, cnt); ++launches;
            CK(cudaMemcpyAsync(xs_r.p, out_r[res].p, (size_t)nvert * D * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        }
        NK(Nccl::get().GroupStart());
        NK(Nccl::get().AllGather(xs_sig32.p, xr_sig32.p, (size_t)cap * (D + 1), ncclInt32, comm, stream));
        NK(Nccl::get().AllGather(xs_r.p, xr_r.p, (size_t)cap * D, ncclFloat64, comm, stream));
        NK(Nccl::get().GroupEnd());
        CK(out_sig[0].ensure((size_t)std::max<long long>(total, 1) * (D + 1))); CK(out_r[0].ensure((size_t)std::max<long long>(total, 1) * D));
        k_unpack_segments<D><<<dim3((unsigned)blocks_for(cap, 128), (unsigned)world), 128, 0, stream>>>(xr_sig32.p, xr_r.p, world, (u32)cap, xc_counts.p,
                                                                                                        out_sig[0].p, out_r[0].p);
        ++launches;
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        st.exchange_bytes = (int64_t)((size_t)cap * ((D + 1) * 4 + D * 8) * (size_t)(world - 1));
        nvert = total; res = 0; staged = false; staged_r = staged_sig64 = staged_sig32 = false; nb_staged32 = false; have_result = true; counts_cached = false;
        nb_total = -1; own_ptr = nullptr;         // the rows are global now: lists are rebuilt from them on request, for every cell
        st.vertices = nvert; st.kernel_launches = launches;
        if (periodic) {
            // the flags of the canonical images are a function of the rows: recomputed for the merged list
            CK(vflags.ensure(std::max<int64_t>(nvert, 1)));
            CK(cudaMemsetAsync(cert.p, 0, sizeof(CertOut), stream));
            if (nvert > 0) {
                k_certify<D><<<blocks_for(nvert, 128), 128, 0, stream>>>(out_sig[res].p, out_r[res].p, (u32)nvert, (long long)n_user, (long long)n,
                                                                        xs_in.p, halo_origin.p, planes.p, pcert, vflags.p, cert.p);
                ++launches;
            }
            k_publish<<<1, 64, 0, stream>>>((const u32*)cert.p, (u32*)h_cert.p, (int)(sizeof(CertOut) / 4), nullptr, nullptr, 0, nullptr, nullptr);
            CK(cudaStreamSynchronize(stream));
            st.unique_vertices = h_cert.p->canonical;
            have_flags = true;
        }
        return HVB_OK;
    }

    int export_device(void* sig, void* r, int64_t cap, int64_t* count) override {
        if (merged || perturbed) { err = "hvb_export_device: not available on a context whose cloud was resolved from non-general position"; return HVB_ESTATE; }
        if (!have_result) { err = "no search result"; return HVB_ESTATE; }
        CK(cudaSetDevice(prm.device));
        if (count) *count = nvert;
        if (cap < nvert) { err = "export buffer too small"; return HVB_EINVAL; }
        if (nvert > 0) {
            CK(cudaMemcpyAsync(sig, out_sig[res].p, (size_t)nvert * (D + 1) * 8, cudaMemcpyDeviceToDevice, stream));
            CK(cudaMemcpyAsync(r, out_r[res].p, (size_t)nvert * D * 8, cudaMemcpyDeviceToDevice, stream));
        }
        CK(cudaStreamSynchronize(stream));
        return HVB_OK;
    }
    // the gathered rows of all ranks are disjoint (ownership rule) and sorted per rank: they become the result as-is
    int adopt_device(const void* sig, const void* r, int64_t count) override {
        CK(cudaSetDevice(prm.device));
        CK(out_sig[0].ensure((size_t)std::max<int64_t>(count, 1) * (D + 1))); CK(out_r[0].ensure((size_t)std::max<int64_t>(count, 1) * D));
        if (count > 0) {
            CK(cudaMemcpyAsync(out_sig[0].p, sig, (size_t)count * (D + 1) * 8, cudaMemcpyDeviceToDevice, stream));
            CK(cudaMemcpyAsync(out_r[0].p, r, (size_t)count * D * 8, cudaMemcpyDeviceToDevice, stream));
        }
        CK(cudaStreamSynchronize(stream));
        nvert = count; res = 0; staged = false; staged_r = staged_sig64 = staged_sig32 = false; nb_staged32 = false; have_result = true;
        nb_total = -1; own_ptr = nullptr;         // the rows are global now: lists are rebuilt from them on request, for every cell
        st.vertices = nvert;
        return HVB_OK;
    }
    // the same for the raw output of a padded all-gather: segment k holds counts[k] valid rows followed by padding
    int adopt_device_padded(const void* sig, const void* r, int nseg, int64_t seg_cap, const int64_t* counts) override {
        CK(cudaSetDevice(prm.device));
        int64_t total = 0;
        for (int k = 0; k < nseg; ++k) { if (counts[k] < 0 || counts[k] > seg_cap) { err = "bad segment count"; return HVB_EINVAL; } total += counts[k]; }
        CK(out_sig[0].ensure((size_t)std::max<int64_t>(total, 1) * (D + 1))); CK(out_r[0].ensure((size_t)std::max<int64_t>(total, 1) * D));
        int64_t at = 0;
        for (int k = 0; k < nseg; ++k) {
            if (counts[k] == 0) continue;
            CK(cudaMemcpyAsync(out_sig[0].p + (size_t)at * (D + 1), (const long long*)sig + (size_t)k * seg_cap * (D + 1),
                               (size_t)counts[k] * (D + 1) * 8, cudaMemcpyDeviceToDevice, stream));
            CK(cudaMemcpyAsync(out_r[0].p + (size_t)at * D, (const double*)r + (size_t)k * seg_cap * D,
                               (size_t)counts[k] * D * 8, cudaMemcpyDeviceToDevice, stream));
            at += counts[k];
        }
        CK(cudaStreamSynchronize(stream));
        nvert = total; res = 0; staged = false; staged_r = staged_sig64 = staged_sig32 = false; nb_staged32 = false; have_result = true;
        nb_total = -1; own_ptr = nullptr;
        st.vertices = nvert;
        return HVB_OK;
    }
    int merge_device(const void* sig, const void* r, int64_t count) override {
        CK(cudaSetDevice(prm.device));
        for (int i = 0; i < 2; ++i) { CK(out_sig[i].ensure((size_t)std::max<int64_t>(count, 1) * (D + 1))); CK(out_r[i].ensure((size_t)std::max<int64_t>(count, 1) * D)); }
        CK(key_top.ensure(std::max<int64_t>(count, 1))); CK(key_hi.ensure(std::max<int64_t>(count, 1))); CK(key_lo.ensure(std::max<int64_t>(count, 1)));
        u64 ts = next_pow2((u64)count * 2 + 16);
        CK(ptab.ensure(ts));
        CK(cudaMemsetAsync(ptab.p, 0, ts * sizeof(u64), stream));
        CK(cudaMemsetAsync(&sc.p->out_count, 0, sizeof(u32), stream));
        int bits = id_bits();
        if (count > 0) {
            k_merge_rows<D><<<blocks_for(count, 128), 128, 0, stream>>>((const long long*)sig, (const double*)r, (u64)count, bits, ptab.p, ts - 1,
                                                                      out_sig[0].p, out_r[0].p, key_top.p, key_hi.p, key_lo.p, &sc.p->out_count);
            ++launches;
        }
        int rc = read_scalars(); if (rc) return rc;
        nvert = h_sc.p->out_count;
        rc = sort_rows((u32)nvert, bits); if (rc) return rc;
        nb_total = -1; own_ptr = nullptr;         // the rows are global now: lists are rebuilt from them on request
        staged = false; staged_r = staged_sig64 = staged_sig32 = false; nb_staged32 = false;   // staged on the first hvb_view_* / hvb_fetch_*
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        st.vertices = nvert; st.kernel_launches = launches;
        have_result = true;
        return HVB_OK;
    }
};

