# runtests.jl -- the reference's own tests of the raycast path, run through the B200 backend.  UNEXECUTED in the build container
# (no Julia there or on its GPU boxes); the Python mirror of every case below runs on the device in tests/test_gpu_*.py.
#
#     julia --project=<HighVoronoi.jl checkout> julia/runtests.jl          (needs a B200 and libhvb200.so, HVB200_LIB to override)
#
# Each testset names the reference test it mirrors; the bar is the reference's own (|sum(volume) - 1| < 1e-3) plus what the
# seam allows on top: the mesh the backend fills equals the mesh the stock SingleThread search fills, signature by signature.
using Test
using HighVoronoi
using StaticArrays
using Random

include(joinpath(@__DIR__, "HighVoronoiB200.jl"))
using .HighVoronoiB200

# every vertex of a mesh once, external ids, sorted -> coordinates
function vertex_dict(mesh, n)
    out = Dict{Vector{Int64}, Vector{Float64}}()
    for i in 1:n
        for (sig, r) in HighVoronoi.vertices_iterator(mesh, i)      # implemented by every mesh type; lists a vertex at each of its cells
            out[sort(collect(sig))] = collect(r)
        end
    end
    return out
end

function search(xs, dom, threading; method = HighVoronoi.RCStandard)
    searcher = HighVoronoi.Raycast(xs; domain = dom, options = RaycastParameter(Float64; method = method, threading = threading))
    mesh = HighVoronoi.cast_mesh(HighVoronoi.DatabaseVertexStorage(), copy(xs))
    HighVoronoi.voronoi(mesh, searcher = searcher, silence = true)
    return mesh
end

function same_mesh(xs, dom; kwargs...)
    n = length(xs)
    a = vertex_dict(search(xs, dom, SingleThread(); kwargs...), n)
    b = vertex_dict(search(xs, dom, B200Thread(); kwargs...), n)
    keys(a) == keys(b) || return false
    for (sig, r) in a
        x0 = xs[sig[1]]
        sum(abs2, r .- b[sig]) <= (1e-10)^2 * sum(abs2, r .- x0) || return false
    end
    return true
end

@testset "HighVoronoiB200" begin
    # the searcher can be constructed at all (raycast-types.jl:426 needs ThreadSafeDict(dict, ::B200Thread))
    @testset "searcher construction" begin
        xs = VoronoiNodes(rand(3, 100))
        @test HighVoronoi.Raycast(xs; domain = cuboid(3, periodic = []), options = RaycastParameter(Float64; threading = B200Thread())) !== nothing
    end

    # test/rcmethods.jl:4-13
    @testset "RaycastMethods" begin
        function test(MM)
            vg1 = VoronoiGeometry(VoronoiNodes(rand(4, 1000)), cuboid(4, periodic = []), vertex_storage = DatabaseVertexStorage(),
                                  integrate = true, integrand = x -> [1.0], integrator = VI_FAST_POLYGON, silence = true,
                                  search_settings = (method = MM, threading = B200Thread()))
            vd1 = VoronoiData(vg1)
            return abs(sum(vd1.bulk_integral)[1] - 1.0) < 0.001
        end
        @test test(RCCombined)
        @test test(RCOriginal)
        @test test(RCNonGeneralFast)
        @test test(RCNonGeneral)
    end

    # test/multithread.jl:3-12 with the backend in the place of MultiThread(1,1)
    @testset "Multithread seam" begin
        vg1 = VoronoiGeometry(VoronoiNodes(rand(4, 1000)), cuboid(4, periodic = []), vertex_storage = DatabaseVertexStorage(),
                              integrate = true, integrand = x -> [1.0], integrator = VI_FAST_POLYGON, silence = true,
                              search_settings = (threading = B200Thread(),))
        @test abs(sum(VoronoiData(vg1).bulk_integral)[1] - 1.0) < 0.001
    end

    # the mesh itself, against the stock search (what tests/test_gpu_parity.py checks against the restated reference)
    @testset "same mesh as SingleThread" begin
        Random.seed!(1)
        for (dim, n) in ((2, 2000), (3, 1000), (4, 400), (5, 150))
            xs = VoronoiNodes(rand(dim, n))
            @test same_mesh(xs, cuboid(dim, periodic = []))
            @test same_mesh(xs, Boundary())                                   # unbounded: rays through pushray!
        end
        xs = VoronoiNodes(rand(4, 500))
        for MM in (RCCombined, RCOriginal, RCNonGeneralFast)
            @test same_mesh(xs, cuboid(4, periodic = []); method = MM)
        end
    end

    # periodic domains go through the reference's own host-side periodisation around voronoi() (domain.jl:175-213)
    @testset "periodic through the host-side halo" begin
        vg = VoronoiGeometry(VoronoiNodes(rand(3, 1000)), cuboid(3, periodic = [1, 2, 3]), integrate = true, integrator = VI_POLYGON,
                             silence = true, search_settings = (threading = B200Thread(),))
        @test abs(sum(VoronoiData(vg).volume) - 1.0) < 0.001
    end

    # refinement: a non-empty mesh reaches the seam, its vertices travel as seed vertices (meshrefine.jl:199-215)
    @testset "refine!" begin
        vg = VoronoiGeometry(VoronoiNodes(rand(3, 500)), cuboid(3, periodic = []), integrate = true, integrator = VI_POLYGON,
                             silence = true, search_settings = (threading = B200Thread(),))
        refine!(vg, VoronoiNodes(rand(3, 50)), silence = true)
        @test abs(sum(VoronoiData(vg).volume) - 1.0) < 0.001
    end

    # test/fraud.jl / test/periodicgrids.jl style input: non-general position, variable-length signatures
    @testset "cubic lattice" begin
        m = 6
        xs = VoronoiNodes(reduce(hcat, [[(i - 0.5) / m, (j - 0.5) / m, (k - 0.5) / m] for i in 1:m for j in 1:m for k in 1:m]))
        vg = VoronoiGeometry(xs, cuboid(3, periodic = []), integrate = true, integrator = VI_POLYGON, silence = true,
                             search_settings = (threading = B200Thread(),))
        @test abs(sum(VoronoiData(vg).volume) - 1.0) < 0.001
    end

    # ConvexHull (test/convexhull.jl): same facets as the stock walk
    @testset "ConvexHull" begin
        xs = VoronoiNodes(rand(3, 2000))
        cv = B200ConvexHull(xs)
        ref = HighVoronoi.ConvexHull(xs)
        @test length(cv) == length(ref)
        @test Set(sort(collect(s)) for (s, _, _) in cv) == Set(sort(collect(ref[i][1])) for i in 1:length(ref))
    end

    release_contexts!()
end
