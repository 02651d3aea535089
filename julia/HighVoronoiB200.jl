# HighVoronoiB200.jl -- the reference-side binding of libhvb200.so.
#
# NOT EXECUTED in the build container or on its GPU boxes: no Julia is installed there.  The same C ABI is exercised by the
# Python ctypes mirror in highvoronoi.jl_b200/ (tests/ call every entry point bound below).  tests/test_abi.py checks that
# `HvbParams` mirrors `struct hvb_params` field by field.
#
# What a HighVoronoi.jl maintainer adds to select the B200 backend through the existing `search_settings` seam -- the
# complete list; each item names the reference line that makes it necessary:
#
#   1. a threading singleton  `B200Thread`                      next to SingleThread / MultiThread (src/HighVoronoi.jl:60-81).
#      `RaycastParameter` is generic in `typeof(threading)` (raycast-types.jl:312-324), so no change there.
#   2. `HighVoronoi.ThreadSafeDict(dict, ::B200Thread) = dict`  `Raycast(xs; options)` builds its `FEIStorage_global` with
#      `ThreadSafeDict(dict, parameters.threading)` (raycast-types.jl:426), which has methods for SingleThread / MultiThread
#      only (threaddict.jl:38-39): without this method the searcher cannot even be constructed.
#   3. `HighVoronoi._voronoi(mesh, TODO, ..., ::B200Thread)`     next to sysvoronoi.jl:41 (SingleThread) and :50 (MultiThread);
#      `voronoi` dispatches on `searcher.parameters.threading` at sysvoronoi.jl:38.
#
# No other dispatch on the threading object is reached from `voronoi()`: `myqht` / `locktype` (queues.jl:22-24) and
# `MultyRaycast` (raycast-types.jl:298-307) are called inside the cell loop `__voronoi` that method 3 replaces;
# `ThreadsafeProgressMeter(..., ::MultiThread)` (progress.jl:15) and `create_multithreads` (parallelmesh.jl:9) likewise;
# `integrate_geo(::MultiThread, ...)` (geometry.jl:203) dispatches on the `integrate` keyword, not on `search_settings`.
#
# Everything above the seam -- VoronoiGeometry / VoronoiData / refine! / substitute! / ConvexHull -- is unchanged:
#
#   using HighVoronoi, HighVoronoiB200
#   VG = VoronoiGeometry(xs, cuboid(3, periodic=[]); search_settings=(threading=B200Thread(),), integrate=false)
#   VG = VoronoiGeometry(xs, cuboid(5, periodic=[]); search_settings=(threading=B200Thread(ngpus=8),), integrate=false)
#
# Periodic domains: at the `voronoi()` seam every plane is a mirror -- the reference periodises on the host around the search
# (Create_Discrete_Domain domain.jl:175-213 calls voronoi() on generators + halo with a plain Boundary), so method 3 never
# sees a periodic plane and works unchanged inside that orchestration.  `periodic_tessellation` below is the one-call
# alternative that runs halo, search and certificate on the device (hvb_create_periodic).

module HighVoronoiB200

using HighVoronoi
using StaticArrays
import HighVoronoi: _voronoi, nodes, AbstractMesh, RaycastIncircleSkip, ThreadSafeDict

export B200Thread, periodic_tessellation, release_contexts!, B200ConvexHull

const LIB = get(ENV, "HVB200_LIB", joinpath(@__DIR__, "..", "highvoronoi.jl_b200", "lib", "libhvb200.so"))

# mirrors `struct hvb_params` (include/hvb200.h)
mutable struct HvbParams
    variance_tol::Cdouble; break_tol::Cdouble; b_nodes_tol::Cdouble; plane_tolerance::Cdouble; ray_tol::Cdouble
    method::Int32; device::Int32; rank::Int32; world::Int32
    fp32_filter::Int32; on_degenerate::Int32; points_per_cell::Int32; seed_stride::Int32; sort_output::Int32
    tile_size::Int32; neighbors::Int32; persistent::Int32
    vertex_capacity::Int64; probe_scale::Cdouble; periodic_margin::Cdouble
    wire32::Int32; decomposition::Int32
    HvbParams() = new()
end

"""
    B200Thread(; device = 0, ngpus = 1, devices = nothing)     one process drives `ngpus` GPUs (hvb_create_multi)
    B200Thread(device, rank, world)                              this process is slab `rank` of `world` (one process per GPU)

The threading singleton of the B200 backend (item 1).  `ngpus > 1` is the analogue of `MultiThread(ngpus, 1)`: contiguous
slabs of the spatially sorted generator order, one per GPU (parallelmesh.jl:52-87).
"""
struct B200Thread
    device::Int32
    rank::Int32
    world::Int32
    ngpus::Int32
    devices::Vector{Int32}
end
B200Thread(; device::Integer = 0, ngpus::Integer = 1, devices = nothing) =
    B200Thread(Int32(device), Int32(0), Int32(1), Int32(devices === nothing ? ngpus : length(devices)),
               devices === nothing ? Int32[] : Int32.(devices))
B200Thread(device::Integer, rank::Integer, world::Integer) = B200Thread(Int32(device), Int32(rank), Int32(world), Int32(1), Int32[])

# item 2 (raycast-types.jl:426, threaddict.jl:38-39): the search runs on the device, the host-side dictionary is never shared
ThreadSafeDict(dict::ADKV, ::B200Thread) where {K, V, ADKV<:AbstractDict{K, V}} = dict

last_error(ctx) = unsafe_string(ccall((:hvb_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
check(rc, ctx=C_NULL) = rc == 0 || error("libhvb200: code $rc: " * last_error(ctx))     # codes become Julia exceptions

# method singletons -> hvb_params.method (raycast-types.jl:244-284)
method_code(::HighVoronoi.Raycast_Original) = Int32(1)
method_code(::HighVoronoi.Raycast_Combined) = Int32(2)
method_code(::HighVoronoi.Raycast_Non_General) = Int32(3)
method_code(::HighVoronoi.Raycast_Original_Safety) = Int32(4)
method_code(::HighVoronoi.Raycast_Non_General_Skip) = Int32(5)
method_code(::HighVoronoi.Raycast_Original_HP) = Int32(6)
method_code(::HighVoronoi.Raycast_Non_General_Asymptotic_General_HP) = Int32(7)
method_code(_) = Int32(0)                                  # RCStandard = RCNonGeneral = RCNonGeneralHP

# ---- the device context lives with the searcher: created at the first voronoi() call, re-targeted with hvb_set_points when
# the same searcher meets another generator set of the same dimension and domain, destroyed by release_contexts!() --------
mutable struct Context
    ptr::Ptr{Cvoid}
    dim::Int
    nplanes::Int
    multi::Bool
end
# weak keys: a searcher that is garbage-collected takes its device context with it (finalizer below), so a session that
# builds many geometries does not accumulate GPU memory
const CONTEXTS = WeakKeyDict{Any, Context}()
function destroy!(c::Context)
    if c.ptr != C_NULL
        ccall((:hvb_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.ptr)
        c.ptr = C_NULL
    end
    return nothing
end
function release_contexts!()
    for (_, c) in CONTEXTS
        destroy!(c)
    end
    empty!(CONTEXTS)
end
atexit(release_contexts!)

function context_for(searcher, xs::Vector{P}, threading::B200Thread) where {P}
    d = size(P)[1]; n = length(xs)
    planes = searcher.domain.planes               # boundary.jl:22-29
    np = length(planes)
    multi = threading.ngpus > 1
    old = get(CONTEXTS, searcher, nothing)
    if old !== nothing && old.ptr != C_NULL && old.dim == d && old.nplanes == np && old.multi == multi
        GC.@preserve xs check(ccall((:hvb_set_points, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), old.ptr, n, pointer(reinterpret(Float64, xs))), old.ptr)
        return old.ptr
    end
    old !== nothing && destroy!(old)
    base = Matrix{Float64}(undef, d, np); normal = Matrix{Float64}(undef, d, np)
    for (k, pl) in enumerate(planes)
        base[:, k] .= pl.base; normal[:, k] .= pl.normal
    end
    prm = HvbParams()
    ccall((:hvb_default_params, LIB), Cvoid, (Ref{HvbParams},), prm)
    par = searcher.parameters                     # raycast-types.jl:136-171
    prm.variance_tol = par.variance_tol; prm.break_tol = par.break_tol; prm.b_nodes_tol = par.b_nodes_tol
    prm.plane_tolerance = par.plane_tolerance; prm.ray_tol = par.ray_tol
    prm.method = method_code(par.method)
    prm.device = threading.device; prm.rank = threading.rank; prm.world = threading.world
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve xs base normal begin
        if multi
            devs = threading.devices
            check(ccall((:hvb_create_multi, LIB), Cint,
                        (Ref{Ptr{Cvoid}}, Cint, Int64, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ref{HvbParams}, Cint, Ptr{Int32}),
                        ctx, d, n, pointer(reinterpret(Float64, xs)), np, base, normal, C_NULL, prm, threading.ngpus,
                        isempty(devs) ? Ptr{Int32}(C_NULL) : pointer(devs)))
        else
            check(ccall((:hvb_create, LIB), Cint,
                        (Ref{Ptr{Cvoid}}, Cint, Int64, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Ref{HvbParams}),
                        ctx, d, n, pointer(reinterpret(Float64, xs)), np, base, normal, prm))
        end
    end
    c = Context(ctx[], d, np, multi)
    CONTEXTS[searcher] = c
    finalizer(_ -> destroy!(c), searcher)          # RaycastIncircleSkip is a mutable struct (raycast-types.jl:372)
    return ctx[]
end

# the vertices a mesh already holds (refinement callers pass a non-empty mesh, meshrefine.jl:199-215): every vertex once.
# `vertices_iterator(mesh, i)` is the one iterator EVERY mesh type that reaches this seam implements -- RefineMesh forwards it
# (meshrefine.jl:58-59) but not `all_vertices_iterator` -- and it lists a vertex at each of its cells, in external numbering; a
# view may permute the ids, so the owner is not "sig[1]" but the smallest generator: the vertex is taken at cell minimum(sig)
# (plane ids n+p are larger than every generator).  The library orders every seed row itself.
function known_vertices(mesh, n::Int, d::Int)
    sigs = Int64[]; rs = Float64[]
    for i in 1:n
        for (sig, r) in HighVoronoi.vertices_iterator(mesh, i)
            minimum(sig) == i || continue
            length(sig) == d + 1 || error("HighVoronoiB200: the mesh holds a non-general vertex (more than dim+1 generators): not supported by the device search")
            append!(sigs, sig); append!(rs, r)
        end
    end
    return reshape(sigs, d + 1, :), reshape(rs, d, :)
end

"""
    _voronoi(mesh, TODO, ..., threading::B200Thread)                                            (item 3)

Drop-in replacement of the cell loop (`__voronoi`, sysvoronoi.jl:152-215): flatten the generators and the boundary
planes, run the search on the GPU(s), replay the result with the reference's own `push!(mesh, sig=>r)`
(abstractmesh.jl:111) and `pushray!` (abstractmesh.jl:191).  `TODO` is `Iter`; vertices the mesh already holds are
passed as seed vertices, so that only NEW vertices come back (hvb_search).
"""
function _voronoi(mesh::AM, TODO, compact, v_offset, silence, iteration_reset, printsearcher,
                  searcher::RaycastIncircleSkip, intro, threading::B200Thread) where {P, AM<:AbstractMesh{P}}
    xs = nodes(mesh)                              # Vector{SVector{d,Float64}} == n x d row-major doubles (voronoinodes.jl:14)
    d = size(P)[1]
    n = length(xs)
    c = context_for(searcher, xs, threading)
    cells = Int64.(collect(TODO))                 # Iter (1-based)
    all_cells = length(cells) == n
    # `iteration_reset` only steers the progress output (sysvoronoi.jl:171-204), it says nothing about the mesh: whether the mesh
    # already holds vertices is read off the mesh itself (n empty iterators for a fresh one)
    ksig, kr = known_vertices(mesh, n, d)
    nk = size(ksig, 2)
    if threading.ngpus > 1 && (!all_cells || nk > 0)
        error("HighVoronoiB200: Iter subsets / meshes that already hold vertices run on one GPU: use B200Thread() for refinement")
    end
    GC.@preserve cells ksig kr begin
        check(ccall((:hvb_search, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Float64}, Int64, Cint),
                    c, all_cells ? Ptr{Int64}(C_NULL) : pointer(cells), all_cells ? 0 : length(cells),
                    nk > 0 ? pointer(ksig) : Ptr{Int64}(C_NULL), nk > 0 ? pointer(kr) : Ptr{Float64}(C_NULL), nk, d + 1), c)
    end
    nv = Ref{Int64}(0); nr = Ref{Int64}(0); ml = Ref{Int64}(0)
    check(ccall((:hvb_counts, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), c, nv, nr, ml), c)
    # rows are copied out (hvb_fetch_vertices): push! keeps the signature vectors, they must not alias library memory
    if ml[] > d + 1
        # non-general position (cubic grids, ...): the backend resolved it (hvb_params.on_degenerate = 2) and returns the
        # reference's variable-length signatures -- one vertex per cospherical set, all its generators (raycast.jl:926-949)
        off = Vector{Int64}(undef, nv[] + 1)
        check(ccall((:hvb_fetch_vertices_var, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}), c, off, C_NULL, C_NULL), c)
        ids = Vector{Int64}(undef, off[end]); r = Matrix{Float64}(undef, d, nv[])
        check(ccall((:hvb_fetch_vertices_var, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}), c, off, ids, r), c)
        for v in 1:nv[]
            push!(mesh, ids[off[v]+1:off[v+1]] => P(view(r, :, v)))                    # abstractmesh.jl:111
        end
    else
        sig = Matrix{Int64}(undef, d + 1, nv[]); r = Matrix{Float64}(undef, d, nv[])
        check(ccall((:hvb_fetch_vertices, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}), c, sig, r), c)
        for v in 1:nv[]
            push!(mesh, sig[:, v] => P(view(r, :, v)))                                 # abstractmesh.jl:111
        end
    end
    if nr[] > 0
        edge = Matrix{Int64}(undef, d, nr[]); rb = Matrix{Float64}(undef, d, nr[]); ru = similar(rb); node = Vector{Int64}(undef, nr[])
        check(ccall((:hvb_fetch_rays, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}), c, edge, rb, ru, node), c)
        for k in 1:nr[]
            HighVoronoi.pushray!(mesh, edge[:, k], P(view(rb, :, k)), P(view(ru, :, k)), node[k])   # abstractmesh.jl:191
        end
    end
    return mesh, searcher
end

"""
    B200ConvexHull(xs; device = 0)

`ConvexHull(xs)` (chull.jl:213-238) on the device (hvb_convex_hull: gift wrapping, the interior of the tessellation is never
computed).  Same surface as the reference's `ConvexHull`: `length(cv)` facets, `cv[i] == (sig, r, u)` with `sig` the `d`
generating nodes (sorted), `r` a point of the facet's hyperplane (the circumcentre of the nodes inside it -- the point the
reference's projection chull.jl:224-232 yields) and `u` the outer unit normal; iterable.
"""
struct B200ConvexHull{P}
    sig::Matrix{Int64}
    r::Matrix{Float64}
    u::Matrix{Float64}
end
function B200ConvexHull(xs::Vector{P}; device::Integer = 0) where {P}
    d = size(P)[1]; n = length(xs)
    prm = HvbParams()
    ccall((:hvb_default_params, LIB), Cvoid, (Ref{HvbParams},), prm)
    prm.device = device
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve xs check(ccall((:hvb_create, LIB), Cint,
                (Ref{Ptr{Cvoid}}, Cint, Int64, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Ref{HvbParams}),
                ctx, d, n, pointer(reinterpret(Float64, xs)), 0, C_NULL, C_NULL, prm))
    c = ctx[]
    try
        check(ccall((:hvb_convex_hull, LIB), Cint, (Ptr{Cvoid},), c), c)
        nv = Ref{Int64}(0); nr = Ref{Int64}(0); ml = Ref{Int64}(0)
        check(ccall((:hvb_counts, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), c, nv, nr, ml), c)
        edge = Matrix{Int64}(undef, d, nr[]); rb = Matrix{Float64}(undef, d, nr[]); ru = similar(rb); node = Vector{Int64}(undef, nr[])
        nr[] > 0 && check(ccall((:hvb_fetch_rays, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}), c, edge, rb, ru, node), c)
        order = sortperm(collect(eachcol(edge)))
        return B200ConvexHull{P}(edge[:, order], rb[:, order], ru[:, order])
    finally
        ccall((:hvb_destroy, LIB), Cvoid, (Ptr{Cvoid},), c)
    end
end
Base.length(c::B200ConvexHull) = size(c.sig, 2)
Base.getindex(c::B200ConvexHull{P}, i::Int) where {P} = (c.sig[:, i], P(view(c.r, :, i)), P(view(c.u, :, i)))
Base.iterate(c::B200ConvexHull, state = 1) = state > length(c) ? nothing : (c[state], state + 1)

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
