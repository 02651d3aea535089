/* c_example.c -- the C ABI of libhvb200.so from plain C: what any host language with an FFI does (the Julia shim
 * julia/HighVoronoiB200.jl makes the same calls through ccall).
 *
 *   gcc -std=c99 -Iinclude examples/c_example.c -Lhighvoronoi.jl_b200/lib -lhvb200 -Wl,-rpath,$PWD/highvoronoi.jl_b200/lib -lm -o c_example
 *   ./c_example [npoints] [dim]          (needs a B200: without a CUDA device hvb_create answers HVB_ENOGPU, there is no CPU path)
 *
 * Tessellates npoints uniform random points in the unit cube (cuboid(dim, periodic=[]), boundary.jl:510-534), prints the
 * vertex count, the first rows, and checks the reference's own known-answer test: the cell volumes add up to the domain
 * (test/rcmethods.jl:8). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "hvb200.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != HVB_OK) {                                                             \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, hvb_last_error(ctx));          \
            if (ctx) hvb_destroy(ctx);                                                   \
            return rc_ == HVB_ENOGPU ? 3 : 1;                                            \
        }                                                                                \
    } while (0)

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 100000;
    const int dim = argc > 2 ? atoi(argv[2]) : 3;
    hvb_ctx* ctx = NULL;
    if (dim < 2 || dim > HVB_MAX_DIM || n <= dim) { fprintf(stderr, "usage: c_example [npoints > dim] [dim 2..6]\n"); return 2; }

    /* generators: n x dim row-major doubles = Vector{SVector{dim,Float64}} (voronoinodes.jl:14) */
    double* xs = (double*)malloc((size_t)n * dim * sizeof(double));
    srand(1);
    for (int64_t i = 0; i < n * dim; ++i) xs[i] = (rand() + 0.5) / ((double)RAND_MAX + 1.0);

    /* the unit cube: plane 2i-1 is the upper face of axis i, plane 2i the lower one (boundary.jl:517-519) */
    const int P = 2 * dim;
    double base[2 * HVB_MAX_DIM * HVB_MAX_DIM] = {0}, normal[2 * HVB_MAX_DIM * HVB_MAX_DIM] = {0};
    for (int i = 0; i < dim; ++i) {
        base[(2 * i) * dim + i] = 1.0;
        normal[(2 * i) * dim + i] = 1.0;
        normal[(2 * i + 1) * dim + i] = -1.0;
    }

    hvb_params prm;
    hvb_default_params(&prm);            /* RaycastParameter(Float64) defaults (raycast-types.jl:312-324) */
    prm.neighbors = 1;

    CHECK(hvb_create(&ctx, dim, n, xs, P, base, normal, &prm));         /* Raycast(xs; domain, options) */
    CHECK(hvb_search(ctx, NULL, 0, NULL, NULL, 0, 0));                   /* voronoi(mesh; searcher)      */

    int64_t nv = 0, nr = 0, msl = 0;
    CHECK(hvb_counts(ctx, &nv, &nr, &msl));
    printf("%s: %lld generators, d = %d: %lld vertices, %lld unbounded edges, signatures of up to %lld generators\n",
           hvb_version(), (long long)n, dim, (long long)nv, (long long)nr, (long long)msl);
    if (msl == dim + 1) {
        int64_t* sig = (int64_t*)malloc((size_t)nv * (dim + 1) * sizeof(int64_t));
        double* r = (double*)malloc((size_t)nv * dim * sizeof(double));
        CHECK(hvb_fetch_vertices(ctx, sig, r));                          /* push!(mesh, sig => r), 1-based ids, plane p = n + p */
        for (int64_t v = 0; v < nv && v < 3; ++v) {
            printf("  sig = [");
            for (int k = 0; k <= dim; ++k) printf("%s%lld", k ? ", " : "", (long long)sig[v * (dim + 1) + k]);
            printf("]  r = (");
            for (int k = 0; k < dim; ++k) printf("%s%.15g", k ? ", " : "", r[v * dim + k]);
            printf(")\n");
        }
        free(sig); free(r);
    }

    int64_t total = 0;
    CHECK(hvb_neighbor_count(ctx, &total));
    printf("neighbour lists: %.2f entries per cell\n", (double)total / (double)n);

    double* vol = (double*)malloc((size_t)n * sizeof(double));
    CHECK(hvb_cell_volumes(ctx, vol));
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) sum += vol[i];
    printf("sum of the cell volumes - 1 = %.3e\n", sum - 1.0);

    hvb_stats_t st;
    CHECK(hvb_stats(ctx, &st));
    printf("device: index %.3f ms, search %.3f ms, finalize %.3f ms; %lld raycasts, %lld kernel launches\n",
           st.ms_build - st.ms_upload, st.ms_search, st.ms_finalize, (long long)st.raycasts, (long long)st.kernel_launches);

    free(vol); free(xs);
    hvb_destroy(ctx);
    return fabs(sum - 1.0) < 1e-9 ? 0 : 1;
}
