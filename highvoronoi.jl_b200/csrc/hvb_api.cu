// hvb_api.cu -- C ABI (include/hvb200.h) of the B200 raycast vertex search.  The per-dimension contexts live in
// hvb_ctx.cuh / hvb_dim.cu; this file holds no device code.  There is no host compute path: if CUDA is unavailable
// every entry fails.
#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "../../include/hvb200.h"
#include "hvb_ctx_base.hpp"

static std::string g_create_error;

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" {

void hvb_default_params(hvb_params* p) {
    memset(p, 0, sizeof(*p));
    p->variance_tol = 1e-15; p->break_tol = 1e-5; p->b_nodes_tol = 1e-7; p->plane_tolerance = 1e-12; p->ray_tol = 1e-12;
    p->method = 0; p->device = 0; p->rank = 0; p->world = 1; p->fp32_filter = 1; p->on_degenerate = 2;
    p->points_per_cell = 0; p->seed_stride = 0; p->sort_output = 1; p->neighbors = 0; p->persistent = 3; p->vertex_capacity = 0; p->probe_scale = 0.0; p->periodic_margin = 0.0; p->wire32 = 0; p->decomposition = 1;
}

int hvb_create(hvb_ctx** out, int dim, int64_t n, const double* xs, int nplanes, const double* plane_base,
               const double* plane_normal, const hvb_params* params) {
    return hvb_create_periodic(out, dim, n, xs, nplanes, plane_base, plane_normal, nullptr, params);
}

// argument checks shared by the create calls; fills `prm`
static int check_create(hvb_ctx** out, int dim, int64_t n, const double* xs, int nplanes, const double* plane_base,
                        const double* plane_normal, const hvb_params* params, hvb_params& prm, int* ndev_out) {
    if (!out) return HVB_EINVAL;
    *out = nullptr;
    if (params) prm = *params; else hvb_default_params(&prm);
    if (dim < 2 || dim > HVB_MAX_DIM) { g_create_error = "dimension must be 2..6"; return HVB_EINVAL; }
    if (n <= dim || !xs) { g_create_error = "There are not enough points to create a Voronoi tessellation"; return HVB_EINVAL; }   // sysvoronoi.jl:25-27
    if (n > 0x7ff00000LL) { g_create_error = "too many generators"; return HVB_EINVAL; }
    if (prm.method < 0 || prm.method > 7) { g_create_error = "unknown raycast method (0..7, raycast-types.jl:244-284)"; return HVB_EINVAL; }
    if (!(prm.variance_tol >= 0) || !(prm.break_tol > 0) || !(prm.b_nodes_tol > 0) || !(prm.plane_tolerance >= 0) || !(prm.ray_tol > 0)) {
        g_create_error = "tolerances must be positive numbers (raycast-types.jl:226-230)"; return HVB_EINVAL;
    }
    if (nplanes < 0 || nplanes > HVB_MAX_PLANES || (nplanes > 0 && (!plane_base || !plane_normal))) { g_create_error = "bad boundary planes"; return HVB_EINVAL; }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(ce) + "); libhvb200 has no host fallback";
        return HVB_ENOGPU;
    }
    *ndev_out = ndev;
    return HVB_OK;
}

int hvb_create_periodic(hvb_ctx** out, int dim, int64_t n, const double* xs, int nplanes, const double* plane_base,
                        const double* plane_normal, const int32_t* plane_bc, const hvb_params* params) {
    hvb_params prm;
    int ndev = 0;
    int rc = check_create(out, dim, n, xs, nplanes, plane_base, plane_normal, params, prm, &ndev);
    if (rc != HVB_OK) return rc;
    if (prm.device < 0 || prm.device >= ndev) { g_create_error = "bad device ordinal"; return HVB_EINVAL; }
    hvb_ctx* c = nullptr;
    switch (dim) {
        case 2: c = hvb_make_ctx_2(); break;
        case 3: c = hvb_make_ctx_3(); break;
        case 4: c = hvb_make_ctx_4(); break;
        case 5: c = hvb_make_ctx_5(); break;
        case 6: c = hvb_make_ctx_6(); break;
    }
    c->dim = dim; c->n = n; c->P = nplanes; c->prm = prm;
    rc = c->init(xs, plane_base, plane_normal, plane_bc);
    if (rc != HVB_OK) { g_create_error = c->err; delete c; return rc; }
    *out = c;
    return HVB_OK;
}

int hvb_create_multi(hvb_ctx** out, int dim, int64_t n, const double* xs, int nplanes, const double* plane_base,
                     const double* plane_normal, const int32_t* plane_bc, const hvb_params* params, int ngpus, const int32_t* devices) {
    hvb_params prm;
    int ndev = 0;
    int rc = check_create(out, dim, n, xs, nplanes, plane_base, plane_normal, params, prm, &ndev);
    if (rc != HVB_OK) return rc;
    if (ngpus < 1 || ngpus > 64 || (!devices && ngpus > ndev)) { g_create_error = "ngpus must be between 1 and the number of visible CUDA devices"; return HVB_EINVAL; }
    for (int k = 0; devices && k < ngpus; ++k) {
        if (devices[k] < 0 || devices[k] >= ndev) { g_create_error = "bad device ordinal"; return HVB_EINVAL; }
    }
    hvb_ctx* c = hvb_make_multi(dim, n, xs, nplanes, plane_base, plane_normal, plane_bc, prm, ngpus, devices, &rc, &g_create_error);
    if (!c) return rc;
    *out = c;
    return HVB_OK;
}

int hvb_comm_unique_id(void* id128) { return id128 ? hvb_nccl_unique_id(id128, &g_create_error) : HVB_EINVAL; }
int hvb_comm_init(hvb_ctx* ctx, const void* id128) { return ctx ? ctx->comm_init(id128) : HVB_EINVAL; }
int hvb_exchange_counts(hvb_ctx* ctx, int64_t* counts) { return ctx ? ctx->exchange_counts(counts) : HVB_EINVAL; }
int hvb_allgather(hvb_ctx* ctx) { return ctx ? ctx->allgather() : HVB_EINVAL; }

int hvb_set_points(hvb_ctx* ctx, int64_t n, const double* xs) { return ctx ? ctx->set_points(n, xs) : HVB_EINVAL; }

int hvb_search(hvb_ctx* ctx, const int64_t* cells, int64_t ncells, const int64_t* seed_sig, const double* seed_r, int64_t nseed, int sig_stride) {
    if (!ctx) return HVB_EINVAL;
    return ctx->search(cells, ncells, seed_sig, seed_r, nseed, sig_stride);
}
int hvb_convex_hull(hvb_ctx* ctx) { return ctx ? ctx->convex_hull(0) : HVB_EINVAL; }
int hvb_convex_hull_via(hvb_ctx* ctx, int method) { return ctx ? ctx->convex_hull(method) : HVB_EINVAL; }
int hvb_counts(hvb_ctx* ctx, int64_t* nvert, int64_t* nrays, int64_t* max_siglen) { return ctx ? ctx->counts(nvert, nrays, max_siglen) : HVB_EINVAL; }
int hvb_fetch_vertices(hvb_ctx* ctx, int64_t* sig, double* r) { return ctx ? ctx->fetch_vertices(sig, r) : HVB_EINVAL; }
int hvb_fetch_vertices_var(hvb_ctx* ctx, int64_t* off, int64_t* ids, double* r) { return ctx ? ctx->fetch_vertices_var(off, ids, r) : HVB_EINVAL; }
int hvb_fetch_vertices_range(hvb_ctx* ctx, int64_t first, int64_t count, int64_t* sig, double* r) { return ctx ? ctx->fetch_vertices_range(first, count, sig, r) : HVB_EINVAL; }
int hvb_view_vertices(hvb_ctx* ctx, const int64_t** sig, const double** r, int64_t* nvert) { return (ctx && sig && r && nvert) ? ctx->view_vertices(sig, r, nvert) : HVB_EINVAL; }
int hvb_view_vertices32(hvb_ctx* ctx, const int32_t** sig, const double** r, int64_t* nvert) { return (ctx && sig && r && nvert) ? ctx->view_vertices32(sig, r, nvert) : HVB_EINVAL; }
int hvb_view_neighbors32(hvb_ctx* ctx, const int64_t** offsets, const int32_t** ids, int64_t* total) { return (ctx && offsets && ids && total) ? ctx->view_neighbors32(offsets, ids, total) : HVB_EINVAL; }
int hvb_fetch_rays(hvb_ctx* ctx, int64_t* edge, double* base, double* dir, int64_t* node) { return ctx ? ctx->fetch_rays(edge, base, dir, node) : HVB_EINVAL; }
int hvb_neighbor_count(hvb_ctx* ctx, int64_t* total) { return (ctx && total) ? ctx->neighbor_count(total) : HVB_EINVAL; }
int hvb_fetch_neighbors(hvb_ctx* ctx, int64_t* offsets, int64_t* ids) { return ctx ? ctx->fetch_neighbors(offsets, ids) : HVB_EINVAL; }
int hvb_view_neighbors(hvb_ctx* ctx, const int64_t** offsets, const int64_t** ids, int64_t* total) { return (ctx && offsets && ids && total) ? ctx->view_neighbors(offsets, ids, total) : HVB_EINVAL; }
int hvb_export_device(hvb_ctx* ctx, void* sig_dev, void* r_dev, int64_t cap, int64_t* count) { return ctx ? ctx->export_device(sig_dev, r_dev, cap, count) : HVB_EINVAL; }
int hvb_merge_device(hvb_ctx* ctx, const void* sig_dev, const void* r_dev, int64_t count) { return ctx ? ctx->merge_device(sig_dev, r_dev, count) : HVB_EINVAL; }
int hvb_adopt_device_padded(hvb_ctx* ctx, const void* sig_dev, const void* r_dev, int nseg, int64_t seg_cap, const int64_t* counts) { return (ctx && counts) ? ctx->adopt_device_padded(sig_dev, r_dev, nseg, seg_cap, counts) : HVB_EINVAL; }
int hvb_adopt_device(hvb_ctx* ctx, const void* sig_dev, const void* r_dev, int64_t count) { return ctx ? ctx->adopt_device(sig_dev, r_dev, count) : HVB_EINVAL; }
int hvb_halo_count(hvb_ctx* ctx, int64_t* nhalo, int32_t* npairs, double* margin) { return ctx ? ctx->halo_count(nhalo, npairs, margin) : HVB_EINVAL; }
int hvb_fetch_halo(hvb_ctx* ctx, int64_t* origin, int32_t* mult, double* xs) { return ctx ? ctx->fetch_halo(origin, mult, xs) : HVB_EINVAL; }
int hvb_fetch_vertex_flags(hvb_ctx* ctx, uint8_t* flags) { return ctx ? ctx->fetch_vertex_flags(flags) : HVB_EINVAL; }
int hvb_fetch_owned(hvb_ctx* ctx, uint8_t* owned) { return ctx ? ctx->fetch_owned(owned) : HVB_EINVAL; }
int hvb_cell_volumes(hvb_ctx* ctx, double* vol) { return ctx ? ctx->cell_volumes(vol) : HVB_EINVAL; }
int hvb_cell_moments(hvb_ctx* ctx, double* vol, double* first, double* second) { return ctx ? ctx->cell_moments(vol, first, second) : HVB_EINVAL; }
int hvb_cell_areas(hvb_ctx* ctx, double* area) { return ctx ? ctx->cell_areas(area) : HVB_EINVAL; }
int hvb_cell_area_moments(hvb_ctx* ctx, double* area, double* first) { return ctx ? ctx->cell_area_moments(area, first) : HVB_EINVAL; }
int hvb_clean_affected(hvb_ctx* ctx, const int64_t* sig, const double* r, int64_t nv, int sig_stride, int64_t first_new, int64_t n_new, uint8_t* keep, uint8_t* affected) {
    return ctx ? ctx->clean_affected(sig, r, nv, sig_stride, first_new, n_new, keep, affected) : HVB_EINVAL;
}
int hvb_stats(hvb_ctx* ctx, hvb_stats_t* out) {
    if (!ctx || !out) return HVB_EINVAL;
    *out = ctx->st;
    return HVB_OK;
}
const char* hvb_last_error(hvb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
void hvb_destroy(hvb_ctx* ctx) { delete ctx; }
const char* hvb_version(void) { return "hvb200 0.2.0 sm_100a"; }

}  // extern "C"
