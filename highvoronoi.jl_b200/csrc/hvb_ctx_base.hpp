// hvb_ctx_base.hpp -- the abstract context behind the opaque `hvb_ctx*` of include/hvb200.h.  One concrete
// Ctx<D> per dimension is compiled in its own translation unit (hvb_dim.cu); hvb_api.cu dispatches through this
// interface only.  Plain C++ (no device code).
#pragma once
#include <stdint.h>
#include <string>
#include "../../include/hvb200.h"

struct hvb_ctx {
    int dim = 0; int64_t n = 0; int P = 0;
    hvb_params prm;
    std::string err;
    hvb_stats_t st;
    virtual ~hvb_ctx() {}
    virtual int init(const double* xs, const double* pbase, const double* pnormal, const int32_t* plane_bc) = 0;
    virtual int halo_count(int64_t* nhalo, int32_t* npairs, double* margin) = 0;
    virtual int fetch_halo(int64_t* origin, int32_t* mult, double* xs) = 0;
    virtual int fetch_vertex_flags(uint8_t* flags) = 0;
    virtual int fetch_owned(uint8_t* owned) = 0;
    virtual int cell_volumes(double* vol) = 0;
    virtual int cell_moments(double* vol, double* first, double* second) = 0;
    virtual int cell_areas(double* area) = 0;
    virtual int cell_area_moments(double* area, double* first) = 0;
    virtual int clean_affected(const int64_t* sig, const double* r, int64_t nv, int stride, int64_t first_new, int64_t n_new, uint8_t* keep, uint8_t* affected) = 0;
    virtual int set_points(int64_t n, const double* xs) = 0;
    virtual int search(const int64_t* cells, int64_t ncells, const int64_t* seed_sig, const double* seed_r, int64_t nseed, int stride) = 0;
    virtual int convex_hull(int method) = 0;
    virtual int counts(int64_t* nv, int64_t* nr, int64_t* msl) = 0;
    virtual int fetch_vertices(int64_t* sig, double* r) = 0;
    virtual int fetch_vertices_var(int64_t* off, int64_t* ids, double* r) = 0;
    virtual int view_vertices(const int64_t** sig, const double** r, int64_t* nv) = 0;
    virtual int view_vertices32(const int32_t** sig, const double** r, int64_t* nv) = 0;
    virtual int view_neighbors32(const int64_t** off, const int32_t** ids, int64_t* total) = 0;
    virtual int fetch_vertices_range(int64_t first, int64_t count, int64_t* sig, double* r) = 0;
    virtual int fetch_rays(int64_t* edge, double* base, double* dir, int64_t* node) = 0;
    virtual int neighbor_count(int64_t* total) = 0;
    virtual int fetch_neighbors(int64_t* off, int64_t* ids) = 0;
    virtual int view_neighbors(const int64_t** off, const int64_t** ids, int64_t* total) = 0;
    virtual int comm_init(const void* id128) = 0;
    virtual int comm_attach(void* nccl_comm) = 0;
    virtual int exchange_counts(int64_t* counts) = 0;
    virtual int allgather() = 0;
    virtual int export_device(void* sig, void* r, int64_t cap, int64_t* count) = 0;
    virtual int merge_device(const void* sig, const void* r, int64_t count) = 0;
    virtual int adopt_device(const void* sig, const void* r, int64_t count) = 0;
    virtual int adopt_device_padded(const void* sig, const void* r, int nseg, int64_t seg_cap, const int64_t* counts) = 0;
};

// factories, one per translation unit hvb_dim.cu (-DHVB_DIM=2..6)
hvb_ctx* hvb_make_ctx_2();
hvb_ctx* hvb_make_ctx_3();
hvb_ctx* hvb_make_ctx_4();
hvb_ctx* hvb_make_ctx_5();
hvb_ctx* hvb_make_ctx_6();

// hvb_multi.cu
hvb_ctx* hvb_make_multi(int dim, int64_t n, const double* xs, int nplanes, const double* plane_base, const double* plane_normal,
                        const int32_t* plane_bc, const hvb_params& prm, int ngpus, const int32_t* devices, int* rc_out, std::string* err_out);
int hvb_nccl_unique_id(void* id128, std::string* err_out);
