# make_reference_fixtures.jl -- pins the oracle against the REAL reference.  TEST INFRASTRUCTURE, UNEXECUTED in the build
# container (no Julia there, and none on the GPU boxes of this build): anyone with Julia >= 1.9 and HighVoronoi.jl 1.4.2
# runs
#
#     julia oracle/make_reference_fixtures.jl tests/golden/ref
#
# and commits the files it writes.  tests/test_oracle.py::test_reference_fixtures picks up every tests/golden/ref/*.txt and
# requires the oracle (and, on a GPU, the CUDA path) to reproduce them: identical sorted signatures, |dr| <= 1e-10 relative.
# Until such files exist the header of oracle/hv_oracle.cpp says "parity unpinned by the reference itself".
#
# Inputs are written out (not re-generated): Julia's and numpy's generators differ.  File format, plain text:
#   line 1: dim n nplanes nvertices
#   n lines: coordinates of the generators (%.17g)
#   nplanes lines: base (dim numbers) then normal (dim numbers) of each plane, reference numbering (boundary.jl:510-534)
#   nvertices lines: dim+1 sorted 1-based ids (plane p = n+p, docs/src/man/short.md:41-42), then dim coordinates (%.17g)
using HighVoronoi
using Random
using Printf

function external_sig(sig, n, lmesh_planes)
    # internal plane ids are typemax(Int64)-p (FVmesh.jl:212-229); external id of plane p is n+p
    return sort([s > n ? n + (typemax(Int64) - s) : s for s in sig])
end

function fixture(path, dim, n, seed; bounded = true)
    Random.seed!(seed)
    xs = VoronoiNodes(rand(dim, n))
    dom = bounded ? cuboid(dim, periodic = []) : Boundary()
    searcher = HighVoronoi.Raycast(xs; domain = dom)                       # default method RCNonGeneralHP, SingleThread
    mesh = HighVoronoi.cast_mesh(HighVoronoi.DatabaseVertexStorage(), copy(xs))
    HighVoronoi.voronoi(mesh, searcher = searcher, silence = true)
    rows = Tuple{Vector{Int64}, Vector{Float64}}[]
    for i in 1:n
        for (sig, r) in HighVoronoi.vertices_iterator(mesh, i)          # lists a vertex at each of its cells: taken at its smallest generator
            minimum(sig) == i && push!(rows, (external_sig(sig, n, length(dom)), collect(r)))
        end
    end
    sort!(rows, by = x -> x[1])
    open(path, "w") do io
        @printf(io, "%d %d %d %d\n", dim, n, length(dom), length(rows))
        for x in xs
            println(io, join((@sprintf("%.17g", v) for v in x), " "))
        end
        for p in dom.planes
            println(io, join((@sprintf("%.17g", v) for v in vcat(collect(p.base), collect(p.normal))), " "))
        end
        for (sig, r) in rows
            println(io, join(sig, " "), " ", join((@sprintf("%.17g", v) for v in r), " "))
        end
    end
    println("wrote $path: $(length(rows)) vertices")
end

function main()
    out = length(ARGS) >= 1 ? ARGS[1] : "tests/golden/ref"
    mkpath(out)
    for (dim, n) in ((2, 400), (3, 300), (4, 150), (5, 80), (6, 40))
        fixture(joinpath(out, "ref_d$(dim)_n$(n)_cube.txt"), dim, n, 100 + dim)
        fixture(joinpath(out, "ref_d$(dim)_n$(n)_free.txt"), dim, n, 200 + dim; bounded = false)
    end
    fixture(joinpath(out, "ref_d3_n1000_cube.txt"), 3, 1000, 1)           # BASELINE.json configs[0]
end

main()
