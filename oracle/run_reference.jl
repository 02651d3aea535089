# run_reference.jl -- times the REAL reference (HighVoronoi.jl) for bench.py's reference arm.  TEST INFRASTRUCTURE.
#
# UNEXECUTED in the build container (no Julia there); bench.py probes for a `julia` binary (baseline/_ref, PATH) and
# runs this script when it finds one, otherwise it times the C++ restatement (oracle/hv_oracle.cpp).
#
# Protocol of the reference's own harness (src/statistics.jl:98-126): uniform rand(dim, N) in the unit cube, domain
# cuboid(dim, periodic=[]), search only: Raycast(xs; domain) + voronoi(), with threading switched on as the docs describe
# (docs/src/man/multithread.md:18-22).  Vertices are counted once (at the cell of their smallest generator; vertices_iterator, abstractmesh.jl:175-182).
#
# usage: julia run_reference.jl <dim> <npoints> <steps> <warmup> <threads>
# prints one JSON line: {"vertices": total over the timed steps, "seconds": total, "threads": t, "julia": version}
using HighVoronoi
using Random

function count_vertices(mesh, n)
    c = 0
    for i in 1:n
        for (sig, _) in HighVoronoi.vertices_iterator(mesh, i)   # lists a vertex at each of its cells: counted at its smallest generator
            minimum(sig) == i && (c += 1)
        end
    end
    return c
end

function one_step(dim, n, seed, nthreads)
    Random.seed!(seed)
    xs = VoronoiNodes(rand(dim, n))
    threading = nthreads > 1 ? HighVoronoi.MultiThread(nthreads, 1) : HighVoronoi.SingleThread()
    t = @elapsed begin
        searcher = HighVoronoi.Raycast(xs; domain = cuboid(dim, periodic = []),
                                       options = HighVoronoi.RaycastParameter(Float64; threading = threading))
        mesh = HighVoronoi.cast_mesh(HighVoronoi.DatabaseVertexStorage(), copy(xs))
        HighVoronoi.voronoi(mesh, searcher = searcher, silence = true)
    end
    return t, count_vertices(mesh, n)
end

function main()
    dim, n, steps, warmup, nthreads = parse.(Int, ARGS[1:5])
    nthreads = min(nthreads, Threads.nthreads(), 8)          # STATUS has 8 slots indexed by threadid() (chull.jl:185-193)
    one_step(dim, min(n, 1000), 0, nthreads)                   # compile
    for i in 1:warmup
        one_step(dim, n, 1000 + i, nthreads)
    end
    T, V = 0.0, 0
    for i in 1:steps
        t, v = one_step(dim, n, i, nthreads)
        T += t; V += v
    end
    println("{\"vertices\": $V, \"seconds\": $T, \"threads\": $nthreads, \"julia\": \"$(VERSION)\"}")
end

main()
