/* hvb200.h -- C ABI of libhvb200.so, the B200 (sm_100a) raycast vertex-search backend.
 *
 * This is the drop-in boundary for ONE path of HighVoronoi.jl: voronoi(mesh; Iter, searcher)
 * (src/sysvoronoi.jl:21), i.e. the edge walk that grows the Voronoi vertex set.  The reference is pure Julia
 * and has no FFI today; the entry points below are what a Julia `ccall` shim binds (julia/HighVoronoiB200.jl,
 * INTEGRATION.md).  Each entry point names the reference interface it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ types, no exceptions, no callbacks cross this boundary;
 *  - generator ids are 1-based Int64 like the reference; boundary plane p (1-based) appears in a vertex
 *    signature as n+p (docs/src/man/short.md:41-42);
 *  - every call returns 0 or a negative HVB_E* code; hvb_last_error() gives the message;
 *  - there is NO CPU fallback: without a usable CUDA device hvb_create fails with HVB_ENOGPU;
 *  - a context is single-caller; distinct contexts are independent.
 *
 * Limits (checked, reported as HVB_EINVAL / HVB_ENOMEM, never silent)
 *  - dimension 2..6; at most HVB_MAX_PLANES boundary planes;
 *  - generators (+ halo generators of a periodic context) + planes < 2^31, and (dim+1) x bits(ids) <= 192 for the
 *    lexicographic row order (only d = 6 with more than 2^27 generators is refused);
 *  - vertices found by ONE context (one GPU's share) < 2^29: a frontier entry carries the vertex index in 29 bits
 *    (5.3e8 vertex records are more than 180 GB of HBM hold); edge-table slots and queue entries are 32-bit indexed;
 *  - int32 views (hvb_view_vertices32 / hvb_view_neighbors32, wire32) need ids < 2^31 - 1;
 *  - the convex hull: facets < 2^27 (facet index in a ridge slot).
 */
#ifndef HVB200_H
#define HVB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HVB_OK            0
#define HVB_EINVAL       -1   /* bad argument (dimension, sizes, points outside the domain, ...) */
#define HVB_ECUDA        -2   /* CUDA runtime error */
#define HVB_ENOGPU       -3   /* no CUDA device: the library never computes on the host */
#define HVB_ENOMEM       -4   /* device allocation failed / capacity exhausted after retries */
#define HVB_EDEGENERATE  -5   /* a vertex with more than dim+1 cospherical generators was met and on_degenerate did not ask to resolve it */
#define HVB_ESTATE       -6   /* call sequence error (fetch before search, ...) */
#define HVB_EINCOMPLETE  -7   /* a descent failed: some cell has no vertex (raycast.jl:54-59 analogue) */
#define HVB_ENCCL        -8   /* NCCL missing / a collective failed / no communicator on a multi-GPU call */

#define HVB_MAX_DIM     6
#define HVB_MAX_PLANES  32

typedef struct hvb_ctx hvb_ctx;

/* Search settings: RaycastParameter (src/raycast-types.jl:312-324) plus backend knobs. */
typedef struct hvb_params {
    /* the five tolerances of the reference, defaults raycast-types.jl:226-230 */
    double variance_tol;      /* 1e-15 : relative variance of the d+1 squared radii above which a vertex counts as
                                 "suboptimal" (raycast.jl:275-277); coordinates always come from the direct solve +
                                 refinement step that replaces the reference's CG correction (raycast.jl:257-261) */
    double break_tol;         /* 1e-5  : above this a vertex is dropped and counted in hvb_stats_t.rejected (raycast.jl:271) */
    double b_nodes_tol;       /* 1e-7  : in the reference only widens the in-range ball that collects cospherical generators
                                 (raycast.jl:870; adjust_boundary_vertex boundary.jl:444 is a no-op).  Here: accepted and
                                 validated (> 0); candidates inside the reference's tie window mark the cloud as being in
                                 non-general position: resolved or reported, see on_degenerate */
    double plane_tolerance;   /* 1e-12 : half-space slack  c = c1 + |c1| * plane_tolerance (raycast.jl:802-804) and smallest
                                 accepted ray parameter t (raycast.jl:887-889) */
    double ray_tol;           /* 1e-12 : used by the reference only inside FastEdgeIterator (edgeiterate.jl:199,435-460), its
                                 enumeration of non-general vertices; this backend resolves or reports such vertices differently
                                 (on_degenerate); accepted and validated */
    int32_t method;           /* 0 RCStandard/RCNonGeneral/RCNonGeneralHP, 1 RCOriginal, 2 RCCombined, 3 RCNonGeneralFast,
                                 4 RCOriginalSafety, 5 RCNonGeneralSkip, 6 RCOriginalHP, 7 RCNonGeneralCutoff
                                 (raycast-types.jl:244-284).  The methods differ in the PROCEDURE that finds the min-t
                                 generator, not in the winner; every value runs the one exact min-t kernel.  Other values:
                                 HVB_EINVAL.  (The RCOriginal family assumes general position; this backend still reports
                                 a non-general vertex instead of picking one winner silently.) */
    int32_t device;           /* CUDA device ordinal */
    /* slab sharding (parallelmesh.jl:52-87): this context explores slab `rank` of `world` contiguous slabs of
       the spatially sorted generator order; rank=0, world=1 explores everything */
    int32_t rank;
    int32_t world;
    int32_t fp32_filter;      /* 1: FP32 candidate filter + FP64 verification (default); 0: FP64 only */
    int32_t on_degenerate;    /* what hvb_search does when it meets a vertex with more than dim+1 cospherical generators
                                 (non-general position: cubic grids, ...; the reference enumerates such vertices with its
                                 FastEdgeIterator, edgeiterate.jl:82-780, and returns ONE vertex listing all generators):
                                 2 (default): resolve it -- the search is repeated on generators moved by a deterministic
                                    offset of 1e-9 of the cloud's extent (general position), coordinates are solved from the
                                    caller's generators, slivers are dropped and rows with equal coordinates merged into one
                                    vertex with the union of the signatures: the reference's result, variable-length rows
                                    (hvb_counts max_siglen > dim+1, hvb_fetch_vertices_var).  Bounded non-periodic domains,
                                    one GPU, searches over all cells without seed vertices; elsewhere as 0;
                                 0: return HVB_EDEGENERATE;
                                 1: count and continue with one arbitrary winner (unsafe: the mesh may be wrong) */
    int32_t points_per_cell;  /* target occupancy of a uniform-grid cell; 0 = auto */
    int32_t seed_stride;      /* one descent seed every `seed_stride` generators; 0 = auto */
    int32_t sort_output;      /* 1: vertices are returned in lexicographic order of their signature (default) */
    int32_t tile_size;        /* lanes cooperating on one frontier entry: 1, 2, 4, 8, 16 or 32; 0 = auto by dimension */
    int32_t neighbors;        /* 1: hvb_search also builds and stages the neighbour lists (default 0: on first request);
                                 with world > 1 they cover the rank's OWN cells (hvb_fetch_owned), other cells get empty lists */
    int32_t persistent;       /* the frontier walk.  3 (default): one persistent launch with a device-side queue; a warp takes
                                 its tickets, vertex indices and queue slots with ONE atomic each, every lane runs the
                                 min-t query of its own ray (tile_size 1 only);
                                 2: the same launch with the warp-cooperative (pooled) query, rows handed out through
                                    shared-memory tickets: correct, measured slower;
                                 4: the pooled query with static scheduling: per round the lanes of a warp run the row
                                    geometry of all its rays' next rows, then the FP32 filter over the resulting chunks of
                                    4 points, 32 chunks per step (hvb_coop.cuh, pool_scan);
                                 1: the persistent launch with one tile per ray from start to end (any tile_size);
                                 0: one launch per frontier round */
    int64_t vertex_capacity;  /* 0 = estimate from lowerbound(d,d) (edgeiteratebase.jl:151); grows on demand */
    double probe_scale;       /* first probe ball radius / circumradius of the origin vertex; 0 = auto */
    double periodic_margin;   /* hvb_create_periodic: first halo margin (distance outside the periodic planes); 0 = auto
                                 from the generator density.  The certificate enlarges it when it proves too small. */
    int32_t wire32;           /* 1: hvb_search stages signatures and neighbour ids in page-locked memory as int32 (read them
                                 with hvb_view_vertices32 / hvb_view_neighbors32): 4 bytes per id less over PCIe.  The int64
                                 calls keep working (they stage the int64 form on first use).  Default 0. */
    int32_t decomposition;    /* world > 1: which cells a rank explores.  0 = contiguous slabs of the spatially sorted order
                                 with equal counts (partition_indices, parallelmesh.jl:52-87: slabs across axis 0);
                                 1 (default) = blocks: the domain is cut along up to three axes (2 x 2 x 2 for 8 ranks) at the
                                 quantiles of the generators' coordinates -- less surface between ranks, hence fewer vertices
                                 found twice, and equal shares of the cheap boundary cells.  Same result either way. */
} hvb_params;

typedef struct hvb_stats_t {
    int64_t vertices;         /* distinct vertices found                                  */
    int64_t rays;             /* unbounded edges (Boundary() without planes)              */
    int64_t raycasts;         /* min-t queries issued (walks + descent steps)             */
    int64_t duplicate_hits;   /* walks that re-found a known vertex                       */
    int64_t closed_skips;     /* frontier entries dropped because the edge was closed     */
    int64_t candidates_fp32;  /* generators run through the FP32 filter                   */
    int64_t candidates_fp64;  /* generators re-evaluated in FP64                          */
    int64_t rows_scanned;     /* grid rows visited                                        */
    int64_t probe_stages;     /* probe balls grown                                        */
    int64_t rounds;           /* frontier rounds                                          */
    int64_t seeds;            /* descents                                                 */
    int64_t degenerate;       /* near-tie winners (non-general position); after on_degenerate = 2 resolved the cloud: vertices with more than dim+1 generators */
    int64_t kernel_launches;  /* kernels of this library launched by the last hvb_search  */
    int64_t capacity_retries;
    double  ms_build;         /* H2D + grid build                                         */
    double  ms_search;        /* seeding + frontier rounds                                */
    double  ms_finalize;      /* canonical coordinates, sort, neighbour lists: result complete in HBM */
    double  ms_expand_kernel; /* device time of the dominant kernel (walk/expand), summed */
    int64_t expand_launches;
    int64_t expand_items;     /* frontier entries processed by it                         */
    double  ms_seed;          /* device time of the first descent kernel                  */
    double  ms_neighbors;     /* neighbour lists (when built inside hvb_search) + staging  */
    double  ms_rows_sort;     /* part of ms_finalize: canonical rows + lexicographic sort  */
    int64_t halo_nodes;       /* periodic contexts: halo generators appended behind the caller's      */
    int64_t unique_vertices;  /* vertices counted once per periodic image class (= vertices otherwise) */
    int64_t periodic_retries; /* searches repeated because the periodic certificate asked for a larger margin */
    double  ms_stage_wait;    /* time hvb_search waited, after the result was complete in HBM (end of ms_finalize), for the
                                 page-locked staging copies (D2H) that overlap the neighbour build */
    double  ms_upload;        /* part of ms_build: the host -> device copy of the generators (hvb_create / hvb_set_points) */
    int64_t rejected;         /* vertices dropped because the relative variance of their squared radii exceeds break_tol
                                 (walkray_correct_vertex raycast.jl:271-273, SRI_vertex_irreparable) */
    int64_t suboptimal;       /* vertices kept with a variance between variance_tol and break_tol (raycast.jl:275-277) */
    int64_t exchange_bytes;   /* bytes this rank received in the last hvb_allgather (0 when the result stayed sharded) */
} hvb_stats_t;

/* fills *p with the reference's defaults (RaycastParameter(Float64), raycast-types.jl:312-324) */
void hvb_default_params(hvb_params* p);

/* Replaces Raycast(xs; domain, options) (src/raycast.jl:26 -> RaycastIncircleSkip raycast-types.jl:412) together
 * with the ExtendedTree / HVKDTree build (extended.jl:92, kd_tree.jl:27-158): copies the generators
 * (n x dim, row-major = Vector{SVector{dim,Float64}}, voronoinodes.jl:14) and the boundary planes
 * (boundary.jl:22-29: base point and outward normal per plane) to the device and builds the spatial index. */
int hvb_create(hvb_ctx** out, int dim, int64_t n, const double* xs,
               int nplanes, const double* plane_base, const double* plane_normal,
               const hvb_params* params);

/* Periodic domains.  Replaces the halo orchestration of VoronoiGeometry for boundaries with periodic planes
 * (Create_Discrete_Domain src/domain.jl:175-213, reflect_nodes :338-390, periodize! :139-166, expand_internal_boundary).
 * plane_bc[p] is Plane.BC of plane p (boundary.jl:15-29): > 0 = 1-based index of the periodic partner plane (parallel,
 * opposite normal; cuboid(d) pairs plane 2i-1 with 2i, boundary.jl:510-534), <= 0 = Dirichlet/Neumann (a mirror on this path).
 * The context appends HALO generators -- copies x + sum_k m_k T_k of caller generators, T_k = one period along
 * the normal of the lower-indexed plane of pair k, m_k integer -- that lie within `margin` outside the periodic
 * planes, pushes those planes outwards by `margin`, and hvb_search explores the cells of the CALLER generators only.
 * After the search a certificate is evaluated on the device: every returned vertex has its empty ball inside the
 * pushed planes, hence no periodic image that was not copied can lie in it and every caller cell is exactly its
 * periodic cell; otherwise the margin grows to what the certificate demands and the search is repeated.
 * Numbering of the results: caller generators 1..n, halo generators n+1..n+nhalo (hvb_fetch_halo), plane p =
 * n+nhalo+p.  Returned are all vertices that touch a caller generator; a vertex whose generators wrap around the
 * domain appears once per image that touches caller generators (each caller cell needs its own image, as in the
 * reference's mesh); hvb_fetch_vertex_flags marks one canonical image per class (hvb_stats_t.unique_vertices).
 * Neighbour lists have n+nhalo+1 offsets; they are built for the caller generators (halo cells get empty lists).
 * plane_bc == NULL is hvb_create.  world > 1 shards a periodic context like any other (the ranks agree on the halo margin by one
 * ncclAllReduce through the context's communicator, hvb_comm_init / hvb_create_multi); seed vertices and hvb_clean_affected are
 * not supported on periodic contexts yet. */
int hvb_create_periodic(hvb_ctx** out, int dim, int64_t n, const double* xs,
                        int nplanes, const double* plane_base, const double* plane_normal, const int32_t* plane_bc,
                        const hvb_params* params);
/* halo of a periodic context (references / reference_shifts of the reference's domain, domain.jl:338-390):
 * origin[i] = 1-based caller generator that halo generator n+1+i copies, mult[i*npairs + k] = m_k, xs = coordinates
 * (nhalo x dim; any output pointer may be NULL).  npairs pairs are numbered by ascending lower plane index. */
int hvb_halo_count(hvb_ctx* ctx, int64_t* nhalo, int32_t* npairs, double* margin);
int hvb_fetch_halo(hvb_ctx* ctx, int64_t* origin, int32_t* mult, double* xs);
/* flags[v] bit 0: vertex v is the canonical image of its periodic class (always 1 on non-periodic contexts) */
int hvb_fetch_vertex_flags(hvb_ctx* ctx, uint8_t* flags);

/* Re-targets an existing context to a new generator set of the same dimension and domain (a second
 * Raycast(xs; domain, options) call in the reference): uploads the points and rebuilds the index, re-using every
 * device and page-locked allocation of the context.  Invalidates the previous result. */
int hvb_set_points(hvb_ctx* ctx, int64_t n, const double* xs);

/* Replaces voronoi(mesh; Iter, searcher) (src/sysvoronoi.jl:21 -> _voronoi :41/:50 -> __voronoi :152):
 * cells = Iter (1-based, NULL = all cells); seed_sig/seed_r = vertices the mesh already holds
 * (nseed rows of sig_stride >= dim+1 ids, 1-based, unused entries 0; the refinement callers
 * meshrefine.jl:199-215 pass a non-empty mesh).  Seed vertices must be general (dim+1 generators); the walk
 * continues from them, only NEW vertices are returned by the fetch calls, the neighbour lists cover both.
 * Blocks until the results are complete on the device. */
int hvb_search(hvb_ctx* ctx, const int64_t* cells, int64_t ncells,
               const int64_t* seed_sig, const double* seed_r, int64_t nseed, int sig_stride);

/* Replaces ConvexHull(xs) = systematic_chull (src/chull.jl:238-387: search_max, descent_chull, a queue of facets whose
 * sub-facets are explored by raycast_des3 :485-499 / peak_direction kd_tree.jl:369-420, explore_chull_vertex :547) on a
 * context created WITHOUT planes.  The interior of the tessellation is never computed: the hull is wrapped facet by facet
 * (csrc/hvb_wrap.cuh): one query per open ridge, the query being a stream of all generators through shared memory (TMA
 * staged) with an FP32 filter and FP64 verification.  Afterwards hvb_counts reports nvert = 0 and nrays = number of facets,
 * and hvb_fetch_rays returns them: edge = the dim generators of the facet (sorted, 1-based), base = the circumcentre of
 * those generators inside the facet's hyperplane (the point the reference stores, chull.jl:224-232), dir = outer unit
 * normal, node = smallest generator; base and dir depend on the facet's generators alone (bitwise reproducible).
 * General position only (a facet with more than dim generators: HVB_EDEGENERATE). */
int hvb_convex_hull(hvb_ctx* ctx);
/* method 0: as hvb_convex_hull; method 1: the walk around the unbounded 2-faces of the Voronoi diagram with the min-t
 * query of the search (csrc/hvb_hull.cuh; base = a point of the unbounded edge the facet is dual to) -- same facets,
 * kept as a cross-check of the wrapping on the Voronoi side, much slower. */
int hvb_convex_hull_via(hvb_ctx* ctx, int method);

/* sizes for the fetch calls; max_siglen is dim+1 unless the cloud was in non-general position and on_degenerate = 2
 * resolved it (then: the largest number of generators of a vertex) */
int hvb_counts(hvb_ctx* ctx, int64_t* nvert, int64_t* nrays, int64_t* max_siglen);

/* Variable-length form of hvb_fetch_vertices, for meshes with vertices of more than dim+1 generators (the reference's
 * sig vectors of non-general vertices, raycast.jl:926-949): vertex v has the sorted 1-based ids ids[off[v] .. off[v+1])
 * (off: nvert+1 entries, ids: off[nvert] entries -- query off first with ids = r = NULL) and coordinates r[v*dim ..].
 * Rows are in lexicographic order.  Works for every result; the fixed-width calls (hvb_fetch_vertices, hvb_view_vertices*)
 * return HVB_ESTATE when max_siglen > dim+1.  On such a mesh the neighbour lists hold the cells that share a FULL
 * interface (neighbors.jl:205-212), and hvb_cell_volumes / hvb_cell_areas / hvb_clean_affected are not available. */
int hvb_fetch_vertices_var(hvb_ctx* ctx, int64_t* off, int64_t* ids, double* r);

/* Replaces the replay target push!(mesh, sig=>r) (src/abstractmesh.jl:111): sig = nvert x (dim+1) sorted
 * 1-based ids, r = nvert x dim. */
int hvb_fetch_vertices(hvb_ctx* ctx, int64_t* sig, double* r);

/* rows [first, first+count) of the same arrays, copied straight from the device: the shard a rank keeps after
 * hvb_merge_device (the whole merged list need not be staged on every rank). */
int hvb_fetch_vertices_range(hvb_ctx* ctx, int64_t first, int64_t count, int64_t* sig, double* r);

/* Replaces pushray!(mesh, full_edge, r, u, _Cell) (src/abstractmesh.jl:191, sysvoronoi.jl:504-511). */
int hvb_fetch_rays(hvb_ctx* ctx, int64_t* edge, double* base, double* dir, int64_t* node);

/* Replaces neighbors_of_cell (src/neighbors.jl:214-262) for every cell: CSR, offsets has n+1 entries. */
int hvb_neighbor_count(hvb_ctx* ctx, int64_t* total);
int hvb_fetch_neighbors(hvb_ctx* ctx, int64_t* offsets, int64_t* ids);

/* Zero-copy variants: pointers into page-locked host memory owned by the context, valid until the next
 * hvb_search / hvb_destroy.  Layout as in hvb_fetch_vertices. */
int hvb_view_vertices(hvb_ctx* ctx, const int64_t** sig, const double** r, int64_t* nvert);
int hvb_view_neighbors(hvb_ctx* ctx, const int64_t** offsets, const int64_t** ids, int64_t* total);
/* The same with ids as int32 (needs n + nplanes < 2^31): the compact wire format.  A caller that wants Int64 widens on
 * its side; the reference's replay loop (push!(mesh, sig => r), abstractmesh.jl:111) copies every signature anyway. */
int hvb_view_vertices32(hvb_ctx* ctx, const int32_t** sig, const double** r, int64_t* nvert);
int hvb_view_neighbors32(hvb_ctx* ctx, const int64_t** offsets, const int32_t** ids, int64_t* total);

/* ---- Multi-GPU (replaces MultiThread(a,b): _voronoi sysvoronoi.jl:50-82, ParallelMesh / partition_indices
 * parallelmesh.jl:52-87, getMultiThreadRaycasters raycast-types.jl:361-371) ------------------------------------------
 * Generators and index are replicated on every GPU; GPU k explores slab k of the spatially sorted generator order and
 * returns the vertices it OWNS (first caller generator in grid order inside the slab): owned sets are disjoint, their
 * union is the full vertex set, so no deduplication traffic is needed.  Two ways to drive it:
 *
 * (1) ONE process, ONE call: hvb_create_multi builds one context per GPU (host thread per GPU, communicators from
 *     ncclCommInitAll) behind a single hvb_ctx.  hvb_search runs the slab searches concurrently; hvb_counts,
 *     hvb_fetch_vertices, hvb_fetch_rays, hvb_fetch_neighbors, hvb_cell_volumes, hvb_stats return the union (rows: the
 *     shards in rank order, each shard sorted; every GPU copies its shard into the caller's buffer over its own PCIe link).
 *     This is what the Julia shim binds for B200Thread(ngpus): the reference's model is one process
 *     (Threads.@threads over slabs, sysvoronoi.jl:74), not one process per device.
 *     devices == NULL: GPUs 0..ngpus-1.  params->rank / world / device are ignored.  (A device listed twice is shared by
 *     two slabs: a way to exercise the decomposition on a one-GPU box; such a context has no communicator.)
 * (2) one process per GPU (MPI / torchrun style): every rank creates its own context with params->rank / world / device,
 *     rank 0 obtains an id with hvb_comm_unique_id and distributes the 128 bytes by whatever channel the host language
 *     has, every rank calls hvb_comm_init.  After hvb_search each rank holds its shard; hvb_exchange_counts tells every
 *     rank all shard sizes (ncclAllGather of one word); hvb_allgather replaces the shard by the rows of ALL ranks
 *     (counts + rows in a compact wire format, (dim+1) int32 + dim doubles per row, one fused NCCL group; rank order,
 *     each shard sorted -- identical on every rank, no dedup needed).  Periodic contexts use the communicator to agree on
 *     the halo margin (ncclAllReduce max of the certificate's demand), so that every rank numbers the halo alike.
 * NCCL is bound at run time (dlopen "libnccl.so.2"); without it these calls return HVB_ENCCL. */
int hvb_create_multi(hvb_ctx** out, int dim, int64_t n, const double* xs,
                     int nplanes, const double* plane_base, const double* plane_normal, const int32_t* plane_bc,
                     const hvb_params* params, int ngpus, const int32_t* devices);
int hvb_comm_unique_id(void* id128 /* out: 128 bytes */);
int hvb_comm_init(hvb_ctx* ctx, const void* id128);
int hvb_exchange_counts(hvb_ctx* ctx, int64_t* counts /* out: world entries */);
int hvb_allgather(hvb_ctx* ctx);

/* Exchange step with the collective issued by the HOST LANGUAGE (kept for callers that bring their own communicator).
 * hvb_export_device copies this rank's vertices (int64 sig rows of dim+1 sorted 1-based caller ids, double r rows)
 * into caller-provided DEVICE buffers of capacity `cap` rows; hvb_merge_device replaces the context's result
 * by the deduplicated, sorted union of `count` gathered rows (device pointers).  The collective itself
 * (NCCL all-gather) is issued by the host language between the two calls. */
int hvb_export_device(hvb_ctx* ctx, void* sig_dev, void* r_dev, int64_t cap, int64_t* count);
int hvb_merge_device(hvb_ctx* ctx, const void* sig_dev, const void* r_dev, int64_t count);
/* With world > 1 a context returns only the vertices it OWNS (first generator in grid order inside its slab): the
 * owned sets of the ranks are disjoint, sorted, and their union is the full vertex set.  After the all-gather the
 * concatenation (rank order) is therefore already deduplicated: hvb_adopt_device installs it as the result without
 * the hash dedup and re-sort of hvb_merge_device. */
int hvb_adopt_device(hvb_ctx* ctx, const void* sig_dev, const void* r_dev, int64_t count);
/* same, for the raw output of a padded all-gather: nseg segments of seg_cap rows, counts[k] (host array) valid rows each */
int hvb_adopt_device_padded(hvb_ctx* ctx, const void* sig_dev, const void* r_dev, int nseg, int64_t seg_cap, const int64_t* counts);

/* owned[i] = 1 (n entries; periodic contexts: the n caller generators) iff cell i+1 belongs to this context's slab
 * (world > 1: its position in the spatially sorted order lies in slab `rank`; the analogue of the index range a thread
 * receives from partition_indices, parallelmesh.jl:52-87).  With world > 1 the neighbour lists are built for owned cells
 * only -- the rank found every vertex of those cells, so they are complete; other cells get EMPTY lists -- until a
 * merged result is installed (hvb_adopt_device*, hvb_merge_device, hvb_allgather), after which they cover every cell. */
int hvb_fetch_owned(hvb_ctx* ctx, uint8_t* owned);

/* Refinement.  Replaces clean_affected! (src/meshrefine.jl:126-149) inside systematic_refine! (:183-216):
 * the context holds ALL generators, old and new (hvb_set_points), the new ones being the id range
 * [first_new, first_new + n_new) (the reference prepends them: first_new = 1); sig/r are the nv vertices of the caller's
 * old mesh in the numbering of the context (sig_stride entries per row, 0 = unused, plane p = n + p).  keep[v] = 1 iff
 * no new generator lies inside the ball of vertex v -- the reference's rule |x_sig1 - r| <= (1 + 1e-7) * dist(r, nearest
 * new node); affected[i] = 1 (n entries) for the new cells and for every generator of a removed vertex.  The new mesh is
 * the surviving rows plus the rows of hvb_search(ctx, new ids, n_new, ...): every vertex that is not an old one names
 * a new generator, so exploring the new cells finds them all (the reference's 1st Voronoi pass, meshrefine.jl:199). */
int hvb_clean_affected(hvb_ctx* ctx, const int64_t* sig, const double* r, int64_t nv, int sig_stride,
                       int64_t first_new, int64_t n_new, uint8_t* keep, uint8_t* affected);

/* Geometry product: volumes of the cells of the caller's generators, computed on the device from the
 * current result rows by the signed flag decomposition of a simple polytope (hvb_geometry.cuh); replaces, for general
 * position, the reference's VI_POLYGON volume pass (integrate.jl:33-53, polyintegrator.jl) behind VoronoiData(...).volume
 * and is the quantity the reference's own tests check (sum of the volumes = volume of the domain, test/rcmethods.jl:8).
 * vol: n doubles (periodic contexts: the n caller generators).  Cells with an unbounded edge get +inf.  Needs every
 * vertex of a cell among the rows: all cells after hvb_search over all cells (after the merge in multi-GPU mode), the
 * cells of Iter otherwise; HVB_ESTATE after a search with seed vertices.  Sums are accumulated in 64-bit fixed point:
 * the result does not depend on the order of the atomics.  A cell one of whose terms leaves the fixed-point range (a vertex
 * far outside the cloud: a bounded cell at the hull of an unbounded domain, nearly parallel facets) gets NaN, not a wrong sum. */
int hvb_cell_volumes(hvb_ctx* ctx, double* vol);
/* Integrals of polynomials up to degree two over every cell, exactly (what VoronoiData(...).bulk_integral holds for
 * such integrands, integrate.jl:33-53 / polyintegrator.jl; the reference's own tests integrate x -> [1, x1^2, x2^2],
 * test/periodicgrids.jl): vol[i] = int 1, first[i*dim + a] = int x_a (centroid = first / vol), second[i*dim*dim + a*dim + b] =
 * int x_a x_b over the cell of generator i.  Any pointer may be NULL.  Every flag of the decomposition behind hvb_cell_volumes
 * is an orthoscheme whose vertices the recursion knows, so the moments of a simplex apply term by term (hvb_geometry.cuh,
 * vertex_flag_moments).  Cells with an unbounded edge: vol = +inf, moments NaN; cells whose sums leave the fixed-point range: NaN.
 * Same completeness rule and fixed-point accumulation as hvb_cell_volumes; single-GPU contexts. */
int hvb_cell_moments(hvb_ctx* ctx, double* vol, double* first, double* second);
/* The same for the interfaces (VoronoiData(...).area): area[k] is the (d-1)-volume of the facet between cell i and
 * ids[k] for every entry k of the CSR neighbour lists of hvb_fetch_neighbors (offsets[i-1] <= k < offsets[i]); the
 * neighbour may be a generator, a halo generator or a boundary plane.  hvb_neighbor_count entries.  Facets that hold an
 * unbounded edge get +inf.  Same completeness rule and the same fixed-point accumulation as hvb_cell_volumes. */
int hvb_cell_areas(hvb_ctx* ctx, double* area);
/* Area and first moment of every interface (VoronoiData(...).interface_integral for integrands up to degree one,
 * integrate.jl:33-53): for entry k of the CSR neighbour lists area[k] = int 1 and first[k*dim + a] = int x_a over the facet between
 * cell i and ids[k] (its centroid = first / area: what a finite-volume flux needs).  The orthoschemes of the flag decomposition
 * restricted to the facet (vertex_flag_moments with `first`).  Either pointer may be NULL.  Unbounded facets: +inf / NaN;
 * entries whose sums leave the fixed-point range: NaN.  Same completeness rule as hvb_cell_areas; single-GPU contexts,
 * general position. */
int hvb_cell_area_moments(hvb_ctx* ctx, double* area, double* first);

/* counters and device times of the last hvb_create / hvb_set_points / hvb_search: the analogue of the searcher's rare_events and
 * of the counters statistics.jl:132-143 reports (raycasts, nn / inrange work per vertex) */
int hvb_stats(hvb_ctx* ctx, hvb_stats_t* out);

const char* hvb_last_error(hvb_ctx* ctx);   /* ctx may be NULL: message of the last failed hvb_create */
void hvb_destroy(hvb_ctx* ctx);

/* library build info: "hvb200 <version> sm_100a" */
const char* hvb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HVB200_H */
