// hvb_multi.cu -- ONE process drives several GPUs behind a single hvb_ctx (hvb_create_multi, include/hvb200.h).
//
// Replaces the reference's MultiThread path (_voronoi(..., ::MultiThread) sysvoronoi.jl:50-82: Threads.@threads over
// contiguous index slabs, every thread with its own Raycast searcher on replicated generators, raycast-types.jl:361-371,
// one shared vertex store behind ParallelMesh / LockMesh, parallelmesh.jl:52-87,250-287).  Here: one host thread and one
// single-GPU context (Ctx<D>, hvb_ctx.cuh) per device, generators and index replicated, slab k of the spatially sorted
// order walked by GPU k, every GPU returning the disjoint set of vertices it owns.  The "shared store" is the caller's
// host buffer: each GPU copies its shard into its segment over its own PCIe link, no row crosses NVLink.  The NCCL
// communicators (ncclCommInitAll) serve the collectives the path does need: the agreement of periodic contexts on the halo
// margin, and hvb_allgather for callers that want the whole list resident on every GPU.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hvb200.h"
#include "hvb_ctx_base.hpp"
#include "hvb_nccl.hpp"

namespace {

struct MultiCtx : hvb_ctx {
    std::vector<hvb_ctx*> sub;
    std::vector<ncclComm_t> comms;
    std::vector<int64_t> cnt_v, cnt_r;          // rows / rays per shard of the last search
    int64_t n_user = 0;
    bool periodic = false;
    bool have_result = false;
    bool gathered = false;                      // after hvb_allgather every GPU holds the whole list: GPU 0 answers alone
    size_t nsrc() const { return gathered ? 1 : sub.size(); }

    ~MultiCtx() override {
        for (hvb_ctx* c : sub) delete c;
        if (hvb::Nccl::get().ok()) for (ncclComm_t c : comms) if (c) hvb::Nccl::get().CommDestroy(c);
    }

    // runs f(k) on one host thread per GPU; the first failure (lowest rank) becomes this context's error
    int each(const std::function<int(int)>& f) {
        const int N = (int)sub.size();
        std::vector<int> rc(N, HVB_OK);
        std::vector<std::thread> th;
        th.reserve(N);
        for (int k = 1; k < N; ++k) th.emplace_back([&, k] { rc[k] = f(k); });
        rc[0] = f(0);
        for (auto& t : th) t.join();
        for (int k = 0; k < N; ++k)
            if (rc[k] != HVB_OK) { err = "GPU " + std::to_string(sub[k]->prm.device) + ": " + sub[k]->err; return rc[k]; }
        return HVB_OK;
    }

    int create(int d, int64_t n_, const double* xs, int P_, const double* pbase, const double* pnormal, const int32_t* plane_bc,
               const hvb_params& p, int ngpus, const int32_t* devices) {
        dim = d; n = n_; P = P_; prm = p; n_user = n_;
        if (plane_bc) for (int q = 0; q < P_; ++q) periodic |= plane_bc[q] > 0;
        sub.assign(ngpus, nullptr);
        std::vector<int> devs(ngpus);
        for (int k = 0; k < ngpus; ++k) devs[k] = devices ? devices[k] : k;
        for (int k = 0; k < ngpus; ++k) {
            hvb_ctx* c = nullptr;
            switch (d) {
                case 2: c = hvb_make_ctx_2(); break;
                case 3: c = hvb_make_ctx_3(); break;
                case 4: c = hvb_make_ctx_4(); break;
                case 5: c = hvb_make_ctx_5(); break;
                default: c = hvb_make_ctx_6(); break;
            }
            c->dim = d; c->n = n_; c->P = P_; c->prm = p;
            c->prm.device = devs[k]; c->prm.rank = k; c->prm.world = ngpus;
            sub[k] = c;
        }
        // communicators first (ncclCommInitAll is one call for all devices), then the per-GPU contexts in parallel
        bool distinct = true;
        for (int k = 0; k < ngpus; ++k) for (int j = 0; j < k; ++j) distinct &= devs[j] != devs[k];
        // (a device listed twice is shared by two slabs -- a way to exercise the decomposition on a small box; NCCL wants
        // distinct devices, so such a context has no communicator: hvb_allgather and periodic margin agreement are refused)
        if (ngpus > 1 && distinct) {
            hvb::Nccl& N = hvb::Nccl::get();
            if (!N.ok()) { err = "NCCL is not available: " + N.error; return HVB_ENCCL; }
            comms.assign(ngpus, nullptr);
            ncclResult_t r = N.CommInitAll(comms.data(), ngpus, devs.data());
            if (r != ncclSuccess) { err = std::string("ncclCommInitAll failed: ") + N.GetErrorString(r); comms.clear(); return HVB_ENCCL; }
            for (int k = 0; k < ngpus; ++k) sub[k]->comm_attach(comms[k]);
        }
        return each([&](int k) { return sub[k]->init(xs, pbase, pnormal, plane_bc); });
    }

    int init(const double*, const double*, const double*, const int32_t*) override { return HVB_ESTATE; }
    int comm_init(const void*) override { err = "a multi-GPU context owns its communicators"; return HVB_ESTATE; }
    int comm_attach(void*) override { return HVB_ESTATE; }

    int set_points(int64_t n_new, const double* xs) override {
        have_result = false; gathered = false; nb_ready = false;
        int rc = each([&](int k) { return sub[k]->set_points(n_new, xs); });
        if (rc == HVB_OK) { n = n_new; n_user = n_new; merge_stats(false); }
        return rc;
    }

    int search(const int64_t* cells, int64_t ncells, const int64_t* seed_sig, const double* seed_r, int64_t nseed, int stride) override {
        have_result = false; gathered = false; nb_ready = false;
        if (cells || nseed > 0) {
            // Iter subsets and seed vertices are the refinement callers (meshrefine.jl:199-215): small, incremental work
            // that one GPU handles; the slab decomposition applies to full searches
            err = "Iter subsets / seed vertices run on a single-GPU context (hvb_create)";
            return HVB_EINVAL;
        }
        int rc = each([&](int k) { return sub[k]->search(nullptr, 0, nullptr, nullptr, 0, 0); });
        if (rc != HVB_OK) return rc;
        cnt_v.assign(sub.size(), 0); cnt_r.assign(sub.size(), 0);
        for (size_t k = 0; k < sub.size(); ++k) { rc = sub[k]->counts(&cnt_v[k], &cnt_r[k], nullptr); if (rc) { err = sub[k]->err; return rc; } }
        merge_stats(true);
        have_result = true;
        return HVB_OK;
    }

    // statistics of the union: counters add up, times are the slowest rank's (the ranks run concurrently)
    void merge_stats(bool searched) {
        hvb_stats_t a = sub[0]->st;
        for (size_t k = 1; k < sub.size(); ++k) {
            const hvb_stats_t& b = sub[k]->st;
            a.ms_build = std::max(a.ms_build, b.ms_build); a.ms_upload = std::max(a.ms_upload, b.ms_upload);
            if (!searched) continue;
            a.vertices += b.vertices; a.rays += b.rays; a.raycasts += b.raycasts; a.duplicate_hits += b.duplicate_hits;
            a.closed_skips += b.closed_skips; a.candidates_fp32 += b.candidates_fp32; a.candidates_fp64 += b.candidates_fp64;
            a.rows_scanned += b.rows_scanned; a.probe_stages += b.probe_stages; a.rounds = std::max(a.rounds, b.rounds);
            a.seeds += b.seeds; a.degenerate += b.degenerate; a.kernel_launches += b.kernel_launches;
            a.capacity_retries = std::max(a.capacity_retries, b.capacity_retries);
            a.ms_search = std::max(a.ms_search, b.ms_search); a.ms_finalize = std::max(a.ms_finalize, b.ms_finalize);
            a.ms_expand_kernel = std::max(a.ms_expand_kernel, b.ms_expand_kernel); a.expand_launches += b.expand_launches;
            a.expand_items += b.expand_items; a.ms_seed = std::max(a.ms_seed, b.ms_seed);
            a.ms_neighbors = std::max(a.ms_neighbors, b.ms_neighbors); a.ms_rows_sort = std::max(a.ms_rows_sort, b.ms_rows_sort);
            a.unique_vertices += b.unique_vertices; a.periodic_retries = std::max(a.periodic_retries, b.periodic_retries);
            a.ms_stage_wait = std::max(a.ms_stage_wait, b.ms_stage_wait); a.rejected += b.rejected; a.suboptimal += b.suboptimal;
            a.exchange_bytes += b.exchange_bytes;
        }
        st = a;
    }

    int need_result() { if (!have_result) { err = "no search result"; return HVB_ESTATE; } return HVB_OK; }

    int counts(int64_t* nv, int64_t* nr, int64_t* msl) override {
        int rc = need_result(); if (rc) return rc;
        int64_t v = 0, r = 0;
        for (size_t k = 0; k < sub.size(); ++k) { v += cnt_v[k]; r += cnt_r[k]; }
        if (nv) *nv = v;
        if (nr) *nr = r;
        if (msl) *msl = dim + 1;
        return HVB_OK;
    }
    int exchange_counts(int64_t* counts_out) override {
        int rc = need_result(); if (rc) return rc;
        if (counts_out) for (size_t k = 0; k < sub.size(); ++k) counts_out[k] = cnt_v[k];
        return HVB_OK;
    }
    // the shards in rank order, each GPU writing its segment of the caller's buffers
    int fetch_vertices(int64_t* sig, double* r) override {
        int rc = need_result(); if (rc) return rc;
        std::vector<int64_t> at(sub.size() + 1, 0);
        for (size_t k = 0; k < sub.size(); ++k) at[k + 1] = at[k] + cnt_v[k];
        return each([&](int k) {
            return sub[k]->fetch_vertices_range(0, cnt_v[k], sig ? sig + at[k] * (dim + 1) : nullptr, r ? r + at[k] * dim : nullptr);
        });
    }
    int fetch_vertices_var(int64_t* off, int64_t* ids, double* r) override {
        int64_t nv = 0; int rc = counts(&nv, nullptr, nullptr); if (rc) return rc;
        if (off) for (int64_t v = 0; v <= nv; ++v) off[v] = v * (dim + 1);
        return fetch_vertices(ids, r);
    }
    int fetch_vertices_range(int64_t first, int64_t count, int64_t* sig, double* r) override {
        int rc = need_result(); if (rc) return rc;
        int64_t tot = 0;
        for (int64_t c : cnt_v) tot += c;
        if (first < 0 || count < 0 || first + count > tot) { err = "row range out of bounds"; return HVB_EINVAL; }
        int64_t at = 0;
        for (size_t k = 0; k < sub.size(); ++k) {
            const int64_t lo = std::max(first, at), hi = std::min(first + count, at + cnt_v[k]);
            if (hi > lo) {
                rc = sub[k]->fetch_vertices_range(lo - at, hi - lo, sig ? sig + (lo - first) * (dim + 1) : nullptr, r ? r + (lo - first) * dim : nullptr);
                if (rc) { err = sub[k]->err; return rc; }
            }
            at += cnt_v[k];
        }
        return HVB_OK;
    }
    int view_vertices(const int64_t**, const double**, int64_t*) override {
        err = "zero-copy views exist per GPU only: use hvb_fetch_vertices on a multi-GPU context";
        return HVB_ESTATE;
    }
    int view_vertices32(const int32_t**, const double**, int64_t*) override {
        err = "zero-copy views exist per GPU only: use hvb_fetch_vertices on a multi-GPU context";
        return HVB_ESTATE;
    }
    int view_neighbors32(const int64_t**, const int32_t**, int64_t*) override {
        err = "zero-copy views exist per GPU only: use hvb_fetch_neighbors on a multi-GPU context";
        return HVB_ESTATE;
    }
    int view_neighbors(const int64_t**, const int64_t**, int64_t*) override {
        err = "zero-copy views exist per GPU only: use hvb_fetch_neighbors on a multi-GPU context";
        return HVB_ESTATE;
    }
    int fetch_rays(int64_t* edge, double* base, double* dir, int64_t* node) override {
        int rc = need_result(); if (rc) return rc;
        std::vector<int64_t> at(sub.size() + 1, 0);
        for (size_t k = 0; k < sub.size(); ++k) at[k + 1] = at[k] + cnt_r[k];
        return each([&](int k) {
            return sub[k]->fetch_rays(edge ? edge + at[k] * dim : nullptr, base ? base + at[k] * dim : nullptr,
                                      dir ? dir + at[k] * dim : nullptr, node ? node + at[k] : nullptr);
        });
    }

    // neighbour lists: every GPU holds the complete lists of the cells it owns and empty lists elsewhere, so the union is
    // a cell-wise concatenation
    std::vector<std::vector<int64_t> > nb_off, nb_ids;
    bool nb_ready = false;
    int gather_neighbors() {
        if (nb_ready) return HVB_OK;
        const size_t N = nsrc();
        nb_off.assign(N, {}); nb_ids.assign(N, {});
        int rc = each([&](int k) {
            if ((size_t)k >= N) return (int)HVB_OK;
            int64_t tot = 0;
            int r2 = sub[k]->neighbor_count(&tot); if (r2) return r2;
            nb_off[k].resize((size_t)sub[k]->n + 1); nb_ids[k].resize((size_t)std::max<int64_t>(tot, 1));
            return sub[k]->fetch_neighbors(nb_off[k].data(), nb_ids[k].data());
        });
        if (rc) return rc;
        nb_ready = true;
        return HVB_OK;
    }
    int neighbor_count(int64_t* total) override {
        int rc = need_result(); if (rc) return rc;
        rc = gather_neighbors(); if (rc) return rc;
        int64_t t = 0;
        for (size_t k = 0; k < nb_off.size(); ++k) t += nb_off[k].back();
        *total = t;
        return HVB_OK;
    }
    int fetch_neighbors(int64_t* off, int64_t* ids) override {
        int rc = need_result(); if (rc) return rc;
        rc = gather_neighbors(); if (rc) return rc;
        const int64_t ncell = (int64_t)nb_off[0].size() - 1;
        int64_t at = 0;
        for (int64_t c = 0; c < ncell; ++c) {
            if (off) off[c] = at;
            for (size_t k = 0; k < nb_off.size(); ++k) {
                const int64_t a = nb_off[k][c], b = nb_off[k][c + 1];
                if (b > a) { if (ids) memcpy(ids + at, nb_ids[k].data() + a, (size_t)(b - a) * sizeof(int64_t)); at += b - a; }
            }
        }
        if (off) off[ncell] = at;
        return HVB_OK;
    }

    int halo_count(int64_t* nhalo, int32_t* npairs, double* margin) override { return fwd(sub[0]->halo_count(nhalo, npairs, margin)); }
    int fetch_halo(int64_t* origin, int32_t* mult, double* xs) override { return fwd(sub[0]->fetch_halo(origin, mult, xs)); }
    int fwd(int rc) { if (rc) err = sub[0]->err; return rc; }
    int fetch_vertex_flags(uint8_t* flags) override {
        int rc = need_result(); if (rc) return rc;
        std::vector<int64_t> at(sub.size() + 1, 0);
        for (size_t k = 0; k < sub.size(); ++k) at[k + 1] = at[k] + cnt_v[k];
        return each([&](int k) { return cnt_v[k] ? sub[k]->fetch_vertex_flags(flags + at[k]) : HVB_OK; });
    }
    int fetch_owned(uint8_t* owned) override {
        // which GPU owns a cell is an internal matter here: the union covers every cell
        memset(owned, 1, (size_t)(periodic ? n_user : n));
        return HVB_OK;
    }
    // a vertex row adds its share to the volume of each of its cells, and every row lives on exactly one GPU: the volume of
    // a cell is the sum of the GPUs' partial volumes
    int cell_volumes(double* vol) override {
        int rc = need_result(); if (rc) return rc;
        const size_t m = (size_t)n_user;
        if (gathered) return fwd(sub[0]->cell_volumes(vol));
        std::vector<std::vector<double> > part(sub.size(), std::vector<double>(m));
        rc = each([&](int k) { return sub[k]->cell_volumes(part[k].data()); });
        if (rc) return rc;
        for (size_t i = 0; i < m; ++i) { double s = 0; for (size_t k = 0; k < sub.size(); ++k) s += part[k][i]; vol[i] = s; }
        return HVB_OK;
    }
    int cell_area_moments(double*, double*) override { err = "hvb_cell_area_moments runs on a single-GPU context (hvb_create)"; return HVB_EINVAL; }
    int cell_moments(double*, double*, double*) override { err = "hvb_cell_moments runs on a single-GPU context (hvb_create)"; return HVB_EINVAL; }
    int cell_areas(double* area) override {
        if (gathered) return fwd(sub[0]->cell_areas(area));
        err = "interface areas need the whole vertex list on one GPU: call hvb_allgather first"; return HVB_ESTATE;
    }
    int clean_affected(const int64_t*, const double*, int64_t, int, int64_t, int64_t, uint8_t*, uint8_t*) override {
        err = "refinement runs on a single-GPU context (hvb_create)"; return HVB_EINVAL;
    }
    int convex_hull(int) override { err = "the convex hull runs on a single-GPU context (hvb_create)"; return HVB_EINVAL; }
    int allgather() override {
        int rc = need_result(); if (rc) return rc;
        rc = each([&](int k) { return sub[k]->allgather(); });
        if (rc) return rc;
        // every GPU now holds the whole list: the union is any one of them
        for (size_t k = 0; k < sub.size(); ++k) { cnt_v[k] = 0; }
        rc = sub[0]->counts(&cnt_v[0], &cnt_r[0], nullptr);
        // rays are not exchanged: they stay sharded (cnt_r unchanged for k > 0)
        for (size_t k = 1; k < sub.size(); ++k) sub[k]->counts(nullptr, &cnt_r[k], nullptr);
        merge_stats(true);
        st.vertices = cnt_v[0];
        nb_ready = false; gathered = true;
        return fwd(rc);
    }
    int export_device(void*, void*, int64_t, int64_t*) override { err = "per-GPU call"; return HVB_ESTATE; }
    int merge_device(const void*, const void*, int64_t) override { err = "per-GPU call"; return HVB_ESTATE; }
    int adopt_device(const void*, const void*, int64_t) override { err = "per-GPU call"; return HVB_ESTATE; }
    int adopt_device_padded(const void*, const void*, int, int64_t, const int64_t*) override { err = "per-GPU call"; return HVB_ESTATE; }
};

}  // namespace

int hvb_nccl_unique_id(void* id128, std::string* err_out) {
    hvb::Nccl& N = hvb::Nccl::get();
    if (!N.ok()) { *err_out = "NCCL is not available: " + N.error; return HVB_ENCCL; }
    ncclUniqueId id;
    ncclResult_t r = N.GetUniqueId(&id);
    if (r != ncclSuccess) { *err_out = std::string("ncclGetUniqueId failed: ") + N.GetErrorString(r); return HVB_ENCCL; }
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, sizeof(id));
    return HVB_OK;
}

hvb_ctx* hvb_make_multi(int dim, int64_t n, const double* xs, int nplanes, const double* plane_base, const double* plane_normal,
                        const int32_t* plane_bc, const hvb_params& prm, int ngpus, const int32_t* devices, int* rc_out, std::string* err_out) {
    MultiCtx* m = new MultiCtx();
    int rc = m->create(dim, n, xs, nplanes, plane_base, plane_normal, plane_bc, prm, ngpus, devices);
    *rc_out = rc;
    if (rc != HVB_OK) { *err_out = m->err; delete m; return nullptr; }
    return m;
}
