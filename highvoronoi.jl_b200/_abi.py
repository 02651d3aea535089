"""ctypes binding of include/hvb200.h.  The library is required: importing this module without a built
libhvb200.so raises, and every compute call fails with HVB_ENOGPU when no CUDA device is present."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HVB_LIB") or os.path.join(_HERE, "lib", "libhvb200.so")   # HVB_LIB: tuning builds

HVB_OK, HVB_EINVAL, HVB_ECUDA, HVB_ENOGPU, HVB_ENOMEM, HVB_EDEGENERATE, HVB_ESTATE, HVB_EINCOMPLETE, HVB_ENCCL = 0, -1, -2, -3, -4, -5, -6, -7, -8
ERROR_NAMES = {-1: "HVB_EINVAL", -2: "HVB_ECUDA", -3: "HVB_ENOGPU", -4: "HVB_ENOMEM", -5: "HVB_EDEGENERATE",
               -6: "HVB_ESTATE", -7: "HVB_EINCOMPLETE", -8: "HVB_ENCCL"}

EXPORTS = ("hvb_default_params", "hvb_create", "hvb_set_points", "hvb_search", "hvb_counts", "hvb_fetch_vertices", "hvb_fetch_vertices_range", "hvb_fetch_rays",
           "hvb_neighbor_count", "hvb_fetch_neighbors", "hvb_view_vertices", "hvb_view_neighbors", "hvb_export_device", "hvb_merge_device", "hvb_adopt_device", "hvb_adopt_device_padded",
           "hvb_stats", "hvb_last_error", "hvb_destroy", "hvb_version", "hvb_create_periodic", "hvb_halo_count", "hvb_fetch_halo", "hvb_fetch_vertex_flags", "hvb_cell_volumes", "hvb_cell_moments", "hvb_cell_areas", "hvb_cell_area_moments", "hvb_cell_area_moments", "hvb_clean_affected", "hvb_fetch_owned",
           "hvb_create_multi", "hvb_comm_unique_id", "hvb_comm_init", "hvb_exchange_counts", "hvb_allgather", "hvb_view_vertices32", "hvb_view_neighbors32", "hvb_convex_hull", "hvb_convex_hull_via", "hvb_fetch_vertices_var")


class hvb_params(ctypes.Structure):
    _fields_ = [("variance_tol", ctypes.c_double), ("break_tol", ctypes.c_double), ("b_nodes_tol", ctypes.c_double),
                ("plane_tolerance", ctypes.c_double), ("ray_tol", ctypes.c_double),
                ("method", ctypes.c_int32), ("device", ctypes.c_int32), ("rank", ctypes.c_int32), ("world", ctypes.c_int32),
                ("fp32_filter", ctypes.c_int32), ("on_degenerate", ctypes.c_int32), ("points_per_cell", ctypes.c_int32),
                ("seed_stride", ctypes.c_int32), ("sort_output", ctypes.c_int32), ("tile_size", ctypes.c_int32), ("neighbors", ctypes.c_int32), ("persistent", ctypes.c_int32),
                ("vertex_capacity", ctypes.c_int64), ("probe_scale", ctypes.c_double), ("periodic_margin", ctypes.c_double),
                ("wire32", ctypes.c_int32), ("decomposition", ctypes.c_int32)]


class hvb_stats_t(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int64) for k in
                ("vertices", "rays", "raycasts", "duplicate_hits", "closed_skips", "candidates_fp32", "candidates_fp64",
                 "rows_scanned", "probe_stages", "rounds", "seeds", "degenerate", "kernel_launches", "capacity_retries")] + \
               [(k, ctypes.c_double) for k in ("ms_build", "ms_search", "ms_finalize", "ms_expand_kernel")] + \
               [("expand_launches", ctypes.c_int64), ("expand_items", ctypes.c_int64)] + \
               [(k, ctypes.c_double) for k in ("ms_seed", "ms_neighbors", "ms_rows_sort")] + \
               [(k, ctypes.c_int64) for k in ("halo_nodes", "unique_vertices", "periodic_retries")] + [("ms_stage_wait", ctypes.c_double), ("ms_upload", ctypes.c_double)] + \
               [("rejected", ctypes.c_int64), ("suboptimal", ctypes.c_int64), ("exchange_bytes", ctypes.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class HVBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERROR_NAMES.get(code, str(code)), msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libhvb200.so is not built (%s); run `python highvoronoi.jl_b200/build.py` -- there is no "
                              "fallback implementation" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        L.hvb_default_params.argtypes = [ctypes.POINTER(hvb_params)]
        L.hvb_default_params.restype = None
        L.hvb_create.argtypes = [ctypes.POINTER(vp), i32, i64, vp, i32, vp, vp, ctypes.POINTER(hvb_params)]
        L.hvb_create_periodic.argtypes = [ctypes.POINTER(vp), i32, i64, vp, i32, vp, vp, vp, ctypes.POINTER(hvb_params)]
        L.hvb_halo_count.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)]
        L.hvb_fetch_halo.argtypes = [vp, vp, vp, vp]
        L.hvb_fetch_vertex_flags.argtypes = [vp, vp]
        L.hvb_cell_volumes.argtypes = [vp, vp]
        L.hvb_cell_moments.argtypes = [vp, vp, vp, vp]
        L.hvb_cell_area_moments.argtypes = [vp, vp, vp]
        L.hvb_fetch_owned.argtypes = [vp, vp]
        L.hvb_create_multi.argtypes = [ctypes.POINTER(vp), i32, i64, vp, i32, vp, vp, vp, ctypes.POINTER(hvb_params), i32, vp]
        L.hvb_comm_unique_id.argtypes = [vp]
        L.hvb_comm_init.argtypes = [vp, vp]
        L.hvb_exchange_counts.argtypes = [vp, vp]
        L.hvb_allgather.argtypes = [vp]
        L.hvb_convex_hull.argtypes = [vp]
        L.hvb_convex_hull_via.argtypes = [vp, ctypes.c_int]
        L.hvb_fetch_vertices_var.argtypes = [vp, vp, vp, vp]
        L.hvb_cell_areas.argtypes = [vp, vp]
        L.hvb_clean_affected.argtypes = [vp, vp, vp, i64, i32, i64, i64, vp, vp]
        L.hvb_set_points.argtypes = [vp, i64, vp]
        L.hvb_search.argtypes = [vp, vp, i64, vp, vp, i64, i32]
        L.hvb_counts.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64)]
        L.hvb_fetch_vertices.argtypes = [vp, vp, vp]
        L.hvb_fetch_vertices_range.argtypes = [vp, i64, i64, vp, vp]
        L.hvb_fetch_rays.argtypes = [vp, vp, vp, vp, vp]
        L.hvb_neighbor_count.argtypes = [vp, ctypes.POINTER(i64)]
        L.hvb_fetch_neighbors.argtypes = [vp, vp, vp]
        L.hvb_view_vertices.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(i64)]
        L.hvb_view_neighbors.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(i64)]
        L.hvb_view_vertices32.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(i64)]
        L.hvb_view_neighbors32.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(i64)]
        L.hvb_export_device.argtypes = [vp, vp, vp, i64, ctypes.POINTER(i64)]
        L.hvb_merge_device.argtypes = [vp, vp, vp, i64]
        L.hvb_adopt_device.argtypes = [vp, vp, vp, i64]
        L.hvb_adopt_device_padded.argtypes = [vp, vp, vp, i32, i64, vp]
        L.hvb_stats.argtypes = [vp, ctypes.POINTER(hvb_stats_t)]
        L.hvb_last_error.argtypes = [vp]
        L.hvb_last_error.restype = ctypes.c_char_p
        L.hvb_destroy.argtypes = [vp]
        L.hvb_destroy.restype = None
        L.hvb_version.restype = ctypes.c_char_p
        for name in ("hvb_convex_hull", "hvb_convex_hull_via", "hvb_fetch_vertices_var", "hvb_view_vertices32", "hvb_view_neighbors32", "hvb_create_multi", "hvb_comm_unique_id", "hvb_comm_init", "hvb_exchange_counts", "hvb_allgather", "hvb_fetch_owned", "hvb_cell_volumes", "hvb_cell_moments", "hvb_cell_areas", "hvb_cell_area_moments", "hvb_clean_affected", "hvb_create", "hvb_create_periodic", "hvb_halo_count", "hvb_fetch_halo", "hvb_fetch_vertex_flags", "hvb_set_points", "hvb_search", "hvb_counts", "hvb_fetch_vertices", "hvb_fetch_vertices_range", "hvb_fetch_rays", "hvb_neighbor_count",
                     "hvb_fetch_neighbors", "hvb_view_vertices", "hvb_view_neighbors", "hvb_export_device", "hvb_merge_device", "hvb_adopt_device", "hvb_adopt_device_padded", "hvb_stats"):
            getattr(L, name).restype = i32
        _lib = L
    return _lib


def check(rc, ctx=None):
    if rc != HVB_OK:
        msg = lib().hvb_last_error(ctx)
        raise HVBError(rc, msg.decode() if msg else "")
