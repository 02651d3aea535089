# HighVoronoiB200.jl -- the reference-side binding of libhvb200.so (NOT executed in the build container: Julia is not
# installed there; the same C ABI is exercised by the Python ctypes mirror in highvoronoi.jl_b200/).
#
# What a HighVoronoi.jl maintainer adds to select the B200 backend through the existing `search_settings` seam:
#
#   1. one new threading singleton  `B200Thread(device, rank, world)`   (next to SingleThread/MultiThread/AutoThread,
#      src/HighVoronoi.jl:60-81)
#   2. one new method of `_voronoi(mesh, TODO, ..., threading::B200Thread)`   (next to src/sysvoronoi.jl:41 and :50)
#
# Everything else -- VoronoiGeometry / VoronoiData / refine! / ConvexHull -- is unchanged: they reach the search only
# through `voronoi(mesh; Iter, searcher)` (sysvoronoi.jl:21), which dispatches on `searcher.parameters.threading`
# (sysvoronoi.jl:38).
#
#   VG = VoronoiGeometry(xs, cuboid(3, periodic=[]); search_settings=(threading=B200Thread(0),), integrate=false)

module HighVoronoiB200

using HighVoronoi
using StaticArrays
import HighVoronoi: _voronoi, nodes, AbstractMesh, RaycastIncircleSkip

const LIB = get(ENV, "HVB200_LIB", joinpath(@__DIR__, "..", "highvoronoi.jl_b200", "lib", "libhvb200.so"))

# mirrors `struct hvb_params` (include/hvb200.h)
mutable struct HvbParams
    variance_tol::Cdouble; break_tol::Cdouble; b_nodes_tol::Cdouble; plane_tolerance::Cdouble; ray_tol::Cdouble
    method::Int32; device::Int32; rank::Int32; world::Int32
    fp32_filter::Int32; on_degenerate::Int32; points_per_cell::Int32; seed_stride::Int32; sort_output::Int32
    tile_size::Int32; neighbors::Int32; persistent::Int32
    vertex_capacity::Int64; probe_scale::Cdouble; periodic_margin::Cdouble
    HvbParams() = new()
end

struct B200Thread            # <: the reference's threading singletons
    device::Int32
    rank::Int32
    world::Int32
end
B200Thread(device::Integer=0) = B200Thread(Int32(device), Int32(0), Int32(1))

last_error(ctx) = unsafe_string(ccall((:hvb_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
check(rc, ctx=C_NULL) = rc == 0 || error("libhvb200: code $rc: " * last_error(ctx))     # codes become Julia exceptions

"""
    _voronoi(mesh, TODO, ..., threading::B200Thread)

Drop-in replacement of the cell loop (`__voronoi`, sysvoronoi.jl:152-215): flatten the generators and the boundary
planes, run the search on the GPU, replay the result with the reference's own `push!(mesh, sig=>r)`
(abstractmesh.jl:111) and `pushray!` (abstractmesh.jl:191).
"""
function _voronoi(mesh::AM, TODO, compact, v_offset, silence, iteration_reset, printsearcher,
                  searcher::RaycastIncircleSkip, intro, threading::B200Thread) where {P, AM<:AbstractMesh{P}}
    xs = nodes(mesh)                              # Vector{SVector{d,Float64}} == n x d row-major doubles (voronoinodes.jl:14)
    d = size(P)[1]
    n = length(xs)
    planes = searcher.domain.planes               # boundary.jl:22-29
    np = length(planes)
    base = Matrix{Float64}(undef, d, np)
    normal = Matrix{Float64}(undef, d, np)
    for (k, pl) in enumerate(planes)
        base[:, k] .= pl.base
        normal[:, k] .= pl.normal
    end
    prm = HvbParams()
    ccall((:hvb_default_params, LIB), Cvoid, (Ref{HvbParams},), prm)
    par = searcher.parameters                     # raycast-types.jl:136-171
    prm.variance_tol = par.variance_tol; prm.break_tol = par.break_tol; prm.b_nodes_tol = par.b_nodes_tol
    prm.plane_tolerance = par.plane_tolerance; prm.ray_tol = par.ray_tol
    prm.device = threading.device; prm.rank = threading.rank; prm.world = threading.world
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve xs base normal begin
        check(ccall((:hvb_create, LIB), Cint,
                    (Ref{Ptr{Cvoid}}, Cint, Int64, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Ref{HvbParams}),
                    ctx, d, n, pointer(reinterpret(Float64, xs)), np, base, normal, prm))
    end
    c = ctx[]
    try
        cells = Int64.(TODO)                      # Iter (1-based)
        all_cells = length(cells) == n
        check(ccall((:hvb_search, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Float64}, Int64, Cint),
                    c, all_cells ? C_NULL : cells, all_cells ? 0 : length(cells), C_NULL, C_NULL, 0, 0), c)
        nv = Ref{Int64}(0); nr = Ref{Int64}(0); ml = Ref{Int64}(0)
        check(ccall((:hvb_counts, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), c, nv, nr, ml), c)
        # zero-copy view of the page-locked result (valid until the context is destroyed)
        psig = Ref{Ptr{Int64}}(C_NULL); pr = Ref{Ptr{Float64}}(C_NULL); cnt = Ref{Int64}(0)
        check(ccall((:hvb_view_vertices, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Int64}}, Ref{Ptr{Float64}}, Ref{Int64}), c, psig, pr, cnt), c)
        sig = unsafe_wrap(Array, psig[], (d + 1, cnt[]))
        r = unsafe_wrap(Array, pr[], (d, cnt[]))
        for v in 1:cnt[]
            push!(mesh, Vector{Int64}(view(sig, :, v)) => P(view(r, :, v)))          # abstractmesh.jl:111
        end
        if nr[] > 0
            edge = Matrix{Int64}(undef, d, nr[]); rb = Matrix{Float64}(undef, d, nr[]); ru = similar(rb); node = Vector{Int64}(undef, nr[])
            check(ccall((:hvb_fetch_rays, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}), c, edge, rb, ru, node), c)
            for k in 1:nr[]
                HighVoronoi.pushray!(mesh, Vector{Int64}(view(edge, :, k)), P(view(rb, :, k)), P(view(ru, :, k)), node[k])   # abstractmesh.jl:191
            end
        end
    finally
        ccall((:hvb_destroy, LIB), Cvoid, (Ptr{Cvoid},), c)
    end
    return mesh, searcher
end

"""
    clean_affected(ctx, sig, r, lnxs, n) -> (keep, affected)

`clean_affected!` (meshrefine.jl:126-149) on the device: `sig` (d+1 x nv, ids of the context that already holds the
prepended new nodes 1..lnxs) and `r` (d x nv) are the old mesh; `keep[v]` tells whether vertex v survives the new
nodes, `affected[i]` marks the new cells and the generators of removed vertices.
"""
function clean_affected(ctx::Ptr{Cvoid}, sig::Matrix{Int64}, r::Matrix{Float64}, lnxs::Integer, n::Integer)
    nv = size(sig, 2)
    keep = Vector{UInt8}(undef, nv); affected = Vector{UInt8}(undef, n)
    check(ccall((:hvb_clean_affected, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}, Int64, Cint, Int64, Int64, Ptr{UInt8}, Ptr{UInt8}),
                ctx, sig, r, nv, size(sig, 1), 1, lnxs, keep, affected), ctx)
    return keep, affected
end

"""
    cell_volumes(ctx, n) -> Vector{Float64}

Volumes of the cells from the vertex rows the context holds (`hvb_cell_volumes`): what `VoronoiData(VG, getvolume=true).volume`
returns after the reference's `VI_POLYGON` pass (integrate.jl:33-53), for general position.
"""
function cell_volumes(ctx::Ptr{Cvoid}, n::Integer)
    vol = Vector{Float64}(undef, n)
    check(ccall((:hvb_cell_volumes, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, vol), ctx)
    return vol
end

"""
    periodic_tessellation(xs, domain; device = 0) -> (sig, r, canonical, origin, mult)

Periodic boundaries (`Plane.BC > 0`, boundary.jl:15-29): the halo orchestration of `VoronoiGeometry`
(Create_Discrete_Domain domain.jl:175-213, reflect_nodes :338-390, periodize! :139-166) runs on the device
(`hvb_create_periodic`: halo copies, pushed planes, certificate with margin retry).  Rows are numbered caller 1..n,
halo n+1..n+nhalo (`origin[k]`, `mult[:, k]` = generator and period multiplicities of halo node k, the reference's
`references` / `reference_shifts`), plane n+nhalo+p; `canonical[v]` marks one image per periodic class.
"""
function periodic_tessellation(xs::Vector{P}, domain; device::Integer = 0) where {P}
    d = size(P)[1]; n = length(xs)
    planes = domain.planes; np = length(planes)
    base = Matrix{Float64}(undef, d, np); normal = Matrix{Float64}(undef, d, np); bc = Vector{Int32}(undef, np)
    for (k, pl) in enumerate(planes)
        base[:, k] .= pl.base; normal[:, k] .= pl.normal; bc[k] = Int32(pl.BC)
    end
    prm = HvbParams(); ccall((:hvb_default_params, LIB), Cvoid, (Ref{HvbParams},), prm); prm.device = device
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve xs base normal bc begin
        check(ccall((:hvb_create_periodic, LIB), Cint,
                    (Ref{Ptr{Cvoid}}, Cint, Int64, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ref{HvbParams}),
                    ctx, d, n, pointer(reinterpret(Float64, xs)), np, base, normal, bc, prm))
    end
    c = ctx[]
    try
        check(ccall((:hvb_search, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Float64}, Int64, Cint),
                    c, C_NULL, 0, C_NULL, C_NULL, 0, 0), c)
        nv = Ref{Int64}(0); nr = Ref{Int64}(0); ml = Ref{Int64}(0)
        check(ccall((:hvb_counts, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), c, nv, nr, ml), c)
        sig = Matrix{Int64}(undef, d + 1, nv[]); r = Matrix{Float64}(undef, d, nv[]); canonical = Vector{UInt8}(undef, nv[])
        check(ccall((:hvb_fetch_vertices, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}), c, sig, r), c)
        check(ccall((:hvb_fetch_vertex_flags, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), c, canonical), c)
        nh = Ref{Int64}(0); npairs = Ref{Int32}(0); margin = Ref{Float64}(0)
        check(ccall((:hvb_halo_count, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int32}, Ref{Float64}), c, nh, npairs, margin), c)
        origin = Vector{Int64}(undef, nh[]); mult = Matrix{Int32}(undef, npairs[], nh[]); hxs = Matrix{Float64}(undef, d, nh[])
        check(ccall((:hvb_fetch_halo, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}), c, origin, mult, hxs), c)
        return sig, r, canonical, origin, mult
    finally
        ccall((:hvb_destroy, LIB), Cvoid, (Ptr{Cvoid},), c)
    end
end

end # module
