"""ctypes binding of include/rbp.h.  Fails loudly when the CUDA extension is missing — never falls back."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RBP_LIB_PATH: another build of the same library (A/B timing of a kernel change); never a fallback
LIB_PATH = os.environ.get("RBP_LIB_PATH") or os.path.join(_HERE, "librbp_b200.so")


class RbpError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        l = load_library()
        msg = l.rbp_status_string(status).decode()
        detail = l.rbp_last_error().decode()
        super().__init__(f"{where}: {msg} ({status}) {detail}")


class Encounter(ctypes.Structure):
    _fields_ = [("weight", ctypes.c_float), ("regret", ctypes.c_float), ("payoff", ctypes.c_float), ("visits", ctypes.c_uint32)]


class ProfileRow(ctypes.Structure):
    _fields_ = [("info_key", ctypes.c_uint32), ("action", ctypes.c_uint32), ("row", Encounter)]


class HyperC(ctypes.Structure):
    _fields_ = [("temperature", ctypes.c_float), ("smoothing", ctypes.c_float), ("curiosity", ctypes.c_float),
                ("prune_threshold", ctypes.c_float), ("prune_explore", ctypes.c_float), ("prune_warmup", ctypes.c_uint32),
                ("regret_min", ctypes.c_float)]


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(librbp_b200 has no CPU fallback)")
    l = ctypes.CDLL(LIB_PATH)
    vp, u64, i32, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32
    P = ctypes.POINTER
    l.rbp_status_string.restype = ctypes.c_char_p
    l.rbp_status_string.argtypes = [i32]
    l.rbp_last_error.restype = ctypes.c_char_p
    l.rbp_kernel_launches.restype = u64
    l.rbp_device_count.restype = i32
    l.rbp_hyper_default.argtypes = [P(HyperC)]
    l.rbp_philox4x32_10.argtypes = [P(u32), P(u32), P(u32)]
    l.rbp_solver_create.argtypes = [i32, i32, i32, i32, i32, i32, u64, P(HyperC), i32, P(vp)]
    l.rbp_solver_set_world.argtypes = [vp, i32, i32]
    l.rbp_solver_set_stream.argtypes = [vp, vp]
    l.rbp_solver_destroy.argtypes = [vp]
    l.rbp_solver_destroy.restype = None
    l.rbp_solver_step.argtypes = [vp, u64]
    l.rbp_solver_step_timed.argtypes = [vp, u64, i32, P(ctypes.c_float), P(ctypes.c_float), P(ctypes.c_float)]
    l.rbp_selftest_div_by_count.argtypes = [u32, u32, P(u64)]
    l.rbp_solver_epochs.argtypes = [vp, P(u64)]
    l.rbp_solver_exploitability.argtypes = [vp, P(ctypes.c_float)]
    l.rbp_solver_counters.argtypes = [vp, P(u64)]
    l.rbp_profile_export.argtypes = [vp, P(ProfileRow), i32, P(i32)]
    l.rbp_profile_import.argtypes = [vp, P(ProfileRow), i32, u64]
    l.rbp_profile_averaged.argtypes = [vp, u32, P(ctypes.c_float), i32, P(i32)]
    l.rbp_solver_game_shape.argtypes = [vp, P(i32)]
    l.rbp_solver_sample.argtypes = [vp]
    l.rbp_solver_delta_buffer.argtypes = [vp, P(vp), P(ctypes.c_size_t)]
    l.rbp_solver_fold_gathered.argtypes = [vp, vp, i32]
    i64 = ctypes.c_int64
    l.rbp_eval_batch.argtypes = [vp, i64, vp]
    l.rbp_river_equity_batch.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    l.rbp_river_equity_device.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    f32p = P(ctypes.c_float)
    l.rbp_kmeans_create.argtypes = [i32, i64, i32, i32, vp, i32, P(vp)]
    l.rbp_kmeans_destroy.argtypes = [vp]
    l.rbp_kmeans_destroy.restype = None
    l.rbp_kmeans_init_pp.argtypes = [vp, u64, vp]
    l.rbp_kmeans_set_centroids.argtypes = [vp, vp]
    l.rbp_kmeans_init_bounds.argtypes = [vp]
    l.rbp_kmeans_step.argtypes = [vp, vp, vp, P(u32)]
    l.rbp_kmeans_step_local.argtypes = [vp]
    l.rbp_kmeans_step_finish.argtypes = [vp, vp, vp, P(u32)]
    l.rbp_kmeans_accumulator.argtypes = [vp, P(vp), P(ctypes.c_size_t)]
    l.rbp_kmeans_counters.argtypes = [vp, P(vp), P(vp)]
    l.rbp_kmeans_stream.argtypes = [vp]
    l.rbp_kmeans_stream.restype = vp
    l.rbp_kmeans_assign.argtypes = [vp, vp, vp]
    l.rbp_kmeans_centroids.argtypes = [vp, vp, vp]
    l.rbp_kmeans_metric.argtypes = [vp, vp]
    l.rbp_kmeans_bounds.argtypes = [vp, vp, vp, vp, vp]
    l.rbp_kmeans_timed.argtypes = [vp, i32, i32, f32p]
    if hasattr(l, "rbp_kmeans_sinkhorn_stats"):
        l.rbp_kmeans_sinkhorn_stats.argtypes = [vp, vp, i32]
    l.rbp_measure_fadd_peak.argtypes = [f32p]
    l.rbp_kmeans_screen.argtypes = [vp, ctypes.c_float]
    l.rbp_kmeans_screen_probe.argtypes = [vp, i64, vp, vp]
    l.rbp_kmeans_set_metric.argtypes = [vp, vp, i32]
    l.rbp_sinkhorn_batch.argtypes = [vp, i32, vp, i32, i32, vp, vp, i64, vp, ctypes.c_float, i32, ctypes.c_float, vp]
    l.rbp_isoset_create.argtypes = [i32, i32, P(vp)]
    l.rbp_isoset_destroy.argtypes = [vp]
    l.rbp_isoset_destroy.restype = None
    l.rbp_isoset_size.argtypes = [vp]
    l.rbp_isoset_size.restype = i64
    l.rbp_isoset_export.argtypes = [vp, i64, i64, vp, vp, vp]
    l.rbp_isoset_set_abstractions.argtypes = [vp, vp]
    l.rbp_isoset_river_buckets.argtypes = [vp]
    l.rbp_isoset_project.argtypes = [vp, vp, i32, i64, i64, vp, P(u64)]
    l.rbp_canonical_batch.argtypes = [vp, vp, i64, vp, vp, vp]
    l.rbp_nlhe_create.argtypes = [i32, i32, i32, i32, u64, P(HyperC), u64, i32, i32, P(vp)]
    l.rbp_nlhe_destroy.argtypes = [vp]
    l.rbp_nlhe_destroy.restype = None
    l.rbp_nlhe_set_world.argtypes = [vp, i32, i32]
    l.rbp_nlhe_set_stream.argtypes = [vp, vp]
    l.rbp_nlhe_step.argtypes = [vp, u64]
    l.rbp_nlhe_step_timed.argtypes = [vp, u64, i32, P(ctypes.c_float)]
    l.rbp_nlhe_counters.argtypes = [vp, P(u64)]
    l.rbp_nlhe_export.argtypes = [vp, vp, u64, P(u64)]
    l.rbp_nlhe_import.argtypes = [vp, vp, u64, u64]
    l.rbp_nlhe_sample.argtypes = [vp]
    l.rbp_nlhe_set_lookup.argtypes = [vp, vp]
    l.rbp_nlhe_set_lookup_rows.argtypes = [vp, vp, vp, i64]
    l.rbp_obs_encode.argtypes = [vp, vp, i64, vp]
    l.rbp_obs_encode.restype = None
    l.rbp_obs_decode.argtypes = [vp, i64, vp, vp]
    l.rbp_obs_decode.restype = None
    l.rbp_isoset_export_rows.argtypes = [vp, i64, i64, vp, vp]
    l.rbp_nlhe_partition_records.argtypes = [vp, P(vp), P(u64)]
    l.rbp_nlhe_touched_rows.argtypes = [vp, P(vp), P(u64), P(i32)]
    l.rbp_nlhe_apply_rows.argtypes = [vp, vp, u64]
    l.rbp_nlhe_records.argtypes = [vp, P(vp), P(u64), P(u64), P(i32)]
    l.rbp_nlhe_fold_records.argtypes = [vp, vp, u64]
    l.rbp_nlhe_debug_tree.argtypes = [vp, i32, vp, i32, P(i32)]
    l.rbp_comm_unique_id.argtypes = [vp]
    l.rbp_comm_init.argtypes = [i32, i32, vp, i32, P(vp)]
    l.rbp_comm_destroy.argtypes = [vp]
    l.rbp_comm_destroy.restype = None
    l.rbp_comm_rank.argtypes = [vp]
    l.rbp_comm_size.argtypes = [vp]
    l.rbp_comm_barrier.argtypes = [vp]
    l.rbp_nlhe_attach_comm.argtypes = [vp, vp]
    l.rbp_solver_spend.argtypes = [vp, ctypes.c_double, P(u64), P(ctypes.c_double)]
    l.rbp_subgame_partition.argtypes = [vp, i32, i32, vp, vp]
    l.rbp_subgame_posterior.argtypes = [i32, P(ProfileRow), i32, i32, i32, i32, vp, i32, vp]
    l.rbp_subgame_create.argtypes = [vp, i32, i32, vp, vp, i32, i32, vp, i32, u64, P(vp)]
    l.rbp_subgame_destroy.argtypes = [vp]
    l.rbp_subgame_entries.argtypes = [i32, i32, i32, vp, i32, i32, vp, i32, vp, vp, vp]
    l.rbp_subgame_destroy.restype = None
    l.rbp_subgame_step.argtypes = [vp, u64]
    l.rbp_subgame_spend.argtypes = [vp, ctypes.c_double, P(u64), P(ctypes.c_double)]
    l.rbp_subgame_info.argtypes = [vp, P(u64), vp, vp]
    l.rbp_subgame_export.argtypes = [vp, i32, P(ProfileRow), i32, P(i32)]
    l.rbp_subgame_averaged.argtypes = [vp, i32, u32, P(ctypes.c_float), i32, P(i32)]
    l.rbp_subgame_harvest.argtypes = [vp, u32, vp, vp, vp, i32, P(i32)]
    l.rbp_nlhe_spend.argtypes = [vp, ctypes.c_double, P(u64), P(ctypes.c_double)]
    l.rbp_kmeans_attach_comm.argtypes = [vp, vp]
    l.rbp_solver_attach_comm.argtypes = [vp, vp]
    l.rbp_nlhe_traffic_counters.argtypes = [vp, P(u64)]
    _lib = l
    return l


def lib():
    return load_library()


def check(status, where):
    if status != 0:
        raise RbpError(status, where)
