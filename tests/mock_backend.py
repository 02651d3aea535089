"""TEST INFRASTRUCTURE: runs bench.py's own arm on a box without a GPU by standing in for `torch` (buffers only) and for the
`hvb200` host API with the CPU restatement of the reference (oracle/) -- so that the Python logic of bench.py (workload sizes,
step loop, JSON line, contract keys) is exercised by the CPU suite.  Nothing here is reachable from the product or from a
normal bench run: it exists only when this file is executed (tests/test_bench_contract.py), and the line it prints is marked
"data": "MOCK".

    python tests/mock_backend.py [bench.py arguments]
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


# ---- torch stand-in: page-locked / device buffers become numpy arrays -----------------------------------------------------
class _T:
    def __init__(self, a):
        self.a = a

    def numpy(self):
        return self.a

    def fill_(self, v):
        return self

    def item(self):
        return self.a.ravel()[0]


def _fake_torch():
    t = types.ModuleType("torch")
    t.float64, t.int64, t.uint8 = np.float64, np.int64, np.uint8
    t.empty = lambda shape, dtype=np.float64, pin_memory=False, device=None: _T(np.empty(1 if device is not None else shape, dtype=dtype))
    t.device = lambda *a: a
    cuda = types.ModuleType("torch.cuda")
    cuda.set_device = lambda i: None
    cuda.synchronize = lambda *a: None
    t.cuda = cuda
    return t, cuda


# ---- hvb200 stand-in: the host API of highvoronoi.jl_b200/api.py over the oracle ---------------------------------------------
def _fake_hvb():
    import hv_oracle
    import hvb200 as real                                  # the real package: domains, parameters, threading objects

    class Searcher:
        def __init__(self, xs, domain=None, options=None, periodic=False):
            self.xs, self.domain, self.parameters = xs, domain, options
            self._ctx = self
            self.n, self.dim = xs.shape
            self.multi = getattr(options, "_multi", None)
            self.periodic = False
            self.o = None

        def set_points(self, xs):
            self.xs, self.o = xs, None
            self.n = xs.shape[0]

        def search(self):
            self.o = hv_oracle.run(self.xs, self.domain.base, self.domain.normal)

        def stats(self):
            keys = ("raycasts", "duplicate_hits", "closed_skips", "candidates_fp32", "candidates_fp64", "rows_scanned", "rounds", "seeds",
                    "ms_build", "ms_upload", "ms_search", "ms_finalize", "ms_seed", "ms_neighbors", "ms_rows_sort", "ms_stage_wait",
                    "capacity_retries", "vertices", "unique_vertices", "halo_nodes", "periodic_retries", "rejected", "suboptimal",
                    "kernel_launches", "ms_expand_kernel", "expand_launches")
            s = dict.fromkeys(keys, 0)
            s.update(ms_build=0.2, ms_upload=0.1, kernel_launches=1)
            if self.o is not None:
                V = len(self.o["sig"])
                ms = 1e-3 * (self.o["stats"]["search_us"] + self.o["stats"]["build_us"])
                s.update(vertices=V, unique_vertices=V, raycasts=V, ms_search=ms, ms_finalize=0.1 * ms, ms_expand_kernel=0.8 * ms,
                         expand_launches=1, kernel_launches=20)
            return s

        def close(self):
            pass

    class Mesh:
        def __init__(self, s, copy=False):
            self.sig, self.r = s.o["sig"], s.o["r"]
            self._nb = (s.o["nb_off"], s.o["nb_ids"])

        def neighbors(self):
            return self._nb

        def number_of_vertices(self):
            return len(self.sig)

    class Lib:
        def hvb_search(self, ctx, *a):
            ctx.search()
            return 0

    m = types.ModuleType("hvb200")
    for k in ("cuboid", "Boundary", "RaycastParameter", "B200Thread", "HVBError"):
        setattr(m, k, getattr(real, k))
    m.Raycast, m.VoronoiMesh = Searcher, Mesh
    abi = types.ModuleType("hvb200._abi")
    abi.lib = lambda: Lib()
    abi.check = lambda rc, ctx=None: None if rc == 0 else (_ for _ in ()).throw(RuntimeError(rc))
    m._abi = abi
    return m


if __name__ == "__main__":
    hvb = _fake_hvb()                                      # imports the real package first, then takes its name
    torch, cuda = _fake_torch()
    sys.modules.update({"torch": torch, "torch.cuda": cuda, "hvb200": hvb, "hvb200._abi": hvb._abi})
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    bench.WORKLOADS.update({"C2": (1500, 3), "C4": (150, 5), "C3": (3000, 2), "D4": (400, 4)})     # the workload NAMES of the line, oracle-sized
    bench.cloud.__globals__["ROOT"] = ROOT
    _dumps = bench.json.dumps
    bench.json = types.SimpleNamespace(**{k: getattr(bench.json, k) for k in ("load", "loads")},
                                       dumps=lambda line, **kw: _dumps(dict(line, data="MOCK") if isinstance(line, dict) and "metric" in line else line, **kw))
    bench.main()
