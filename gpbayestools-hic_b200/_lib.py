"""ctypes binding of libgpbt_b200.so (C ABI: include/gpbt.h).

There is deliberately no fallback: if the CUDA library has not been built, importing this module
raises, and every product entry point that needs the GPU fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (GPBT_B200_LIB: another build of the same library, e.g. a tuning variant; read once, at import)
LIB_PATH = os.environ.get("GPBT_B200_LIB") or os.path.join(_HERE, "libgpbt_b200.so")

KERNEL_RBF, KERNEL_MATERN32, KERNEL_PCGP = 0, 1, 2
FLAG_NO_PCA, FLAG_EXP_DIAG = 1, 2
PATH_AUTO, PATH_DENSE, PATH_LOWRANK, PATH_DIAG = 0, 1, 2, 3

# every symbol include/gpbt.h declares (tests/test_cabi.py checks the library exports them all)
SYMBOLS = [
    "gpbt_last_error", "gpbt_version", "gpbt_emulator_create", "gpbt_emulator_destroy",
    "gpbt_emulator_create_pcgp", "gpbt_emulator_set_param_trafo", "gpbt_emulator_input_dim",
    "gpbt_pc_predict", "gpbt_backtransform", "gpbt_backtransform_diag", "gpbt_mvn_loglike", "gpbt_chain_create",
    "gpbt_chain_destroy", "gpbt_chain_predict", "gpbt_log_posterior", "gpbt_log_posterior_scatter",
    "gpbt_log_posterior_host",
    "gpbt_chain_workspace_bytes", "gpbt_launch_count", "gpbt_debug_exp_neg",
    "gpbt_host_temp_exchange", "gpbt_ensemble_create", "gpbt_ensemble_destroy", "gpbt_ensemble_set_state", "gpbt_ensemble_get_state",
    "gpbt_ensemble_run", "gpbt_ensemble_steps", "gpbt_ensemble_reserve",
    "gpbt_ensemble_prepare", "gpbt_ensemble_begin_half", "gpbt_ensemble_copy_proposals", "gpbt_ensemble_end_half", "gpbt_ensemble_read", "gpbt_ensemble_reset",
    "gpbt_device_count", "gpbt_set_device", "gpbt_get_device", "gpbt_set_option",
    "gpbt_debug_timing_read", "gpbt_debug_fused_read", "gpbt_fanout_create", "gpbt_fanout_destroy", "gpbt_fanout_size", "gpbt_fanout_log_posterior_host",
    "gpbt_ptlmc_create", "gpbt_ptlmc_destroy", "gpbt_ptlmc_set_state", "gpbt_ptlmc_run", "gpbt_ptlmc_read",
]


class GpbtError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU implementation to fall back to." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    dp, vp, i64, i32, dbl = C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double
    lib.gpbt_last_error.restype = C.c_char_p
    lib.gpbt_version.restype = i32
    lib.gpbt_launch_count.restype = i64
    lib.gpbt_emulator_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i32, i32] + [dp] * 10
    lib.gpbt_emulator_destroy.argtypes = [vp]
    lib.gpbt_emulator_create_pcgp.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i32] + [dp] * 10
    lib.gpbt_emulator_set_param_trafo.argtypes = [vp, i32, dp, i32, i32, dp, dp, dp, dp, dp, dp, dp, dp]
    lib.gpbt_emulator_input_dim.argtypes = [vp]
    lib.gpbt_pc_predict.argtypes = [vp, dp, dp, dp, dp, i64, i64, vp]
    lib.gpbt_backtransform.argtypes = [vp, dp, dp, i64, dp, i64, dp, i64, i64, i64, vp]
    lib.gpbt_backtransform_diag.argtypes = [vp, dp, dp, i64, dp, dp, i64, i64, i64, vp]
    lib.gpbt_mvn_loglike.argtypes = [dp, dp, dp, dp, dp, dp, dbl, i64, i32, vp]
    lib.gpbt_chain_create.argtypes = [C.POINTER(vp), C.POINTER(vp), i32, i32, dp, dp, dp, dp, dp, dp, dbl, dbl]
    lib.gpbt_chain_destroy.argtypes = [vp]
    lib.gpbt_chain_predict.argtypes = [vp, dp, dbl, dp, dp, i64, vp]
    lib.gpbt_log_posterior.argtypes = [vp, dp, dbl, dp, dp, i64, i32, vp]
    lib.gpbt_log_posterior_scatter.argtypes = [vp, dp, dbl, dp, dp, i32, i64, dp, i64, i32, vp]
    lib.gpbt_log_posterior_host.argtypes = [vp, dp, dbl, dp, dp, i64, i32]
    lib.gpbt_debug_exp_neg.argtypes = [dp, dp, i64, vp]
    lib.gpbt_chain_workspace_bytes.argtypes = [vp]
    lib.gpbt_chain_workspace_bytes.restype = i64
    lib.gpbt_host_temp_exchange.argtypes = [dp, dp, i64, dp, dp, i64, dp]
    lib.gpbt_ensemble_create.argtypes = [C.POINTER(vp), vp, i32, dbl, i32, C.c_uint64]
    lib.gpbt_ensemble_destroy.argtypes = [vp]
    lib.gpbt_ensemble_set_state.argtypes = [vp, dp, dp]
    lib.gpbt_ensemble_get_state.argtypes = [vp, dp, dp]
    lib.gpbt_ensemble_run.argtypes = [vp, i64, dp, dp, dp, i32]
    lib.gpbt_ensemble_steps.argtypes = [vp]
    lib.gpbt_ensemble_steps.restype = i64
    lib.gpbt_ensemble_reserve.argtypes = [vp, i64]
    lib.gpbt_ensemble_prepare.argtypes = [vp, i64]
    lib.gpbt_ensemble_begin_half.argtypes = [vp, i32, vp]
    lib.gpbt_ensemble_copy_proposals.argtypes = [vp, i32, i64, i64, dp, vp]
    lib.gpbt_ensemble_end_half.argtypes = [vp, i32, dp, vp]
    lib.gpbt_ensemble_read.argtypes = [vp, i64, i64, dp, dp, dp, dp]
    lib.gpbt_ensemble_reset.argtypes = [vp]
    lib.gpbt_debug_timing_read.argtypes = [vp, i64]
    lib.gpbt_debug_fused_read.argtypes = [vp, i32, i64, dp, i64]
    lib.gpbt_device_count.argtypes = []
    lib.gpbt_set_device.argtypes = [i32]
    lib.gpbt_get_device.argtypes = []
    lib.gpbt_set_option.argtypes = [C.c_char_p, C.c_char_p]
    lib.gpbt_fanout_create.argtypes = [C.POINTER(vp), C.POINTER(vp), i32]
    lib.gpbt_fanout_destroy.argtypes = [vp]
    lib.gpbt_fanout_size.argtypes = [vp]
    lib.gpbt_fanout_log_posterior_host.argtypes = [vp, dp, dbl, dp, C.POINTER(i32), i64, i32, i32, C.POINTER(i32)]
    lib.gpbt_ptlmc_create.argtypes = [C.POINTER(vp), vp, i32, i32, dp, dp, dbl, C.c_uint64]
    lib.gpbt_ptlmc_destroy.argtypes = [vp]
    lib.gpbt_ptlmc_set_state.argtypes = [vp, dp, dbl]
    lib.gpbt_ptlmc_run.argtypes = [vp, i64, i64, i64]
    lib.gpbt_ptlmc_read.argtypes = [vp, dp, dp, dp, dp]
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise GpbtError("gpbt error %d: %s" % (rc, lib.gpbt_last_error().decode()))


def set_option(key, value=None):
    """gpbt_set_option: tuning override ("pc_tile", "chol", "lowrank_generic", "no_zerocopy",
    "ensemble_split_kernels", "fanout_min_rows", "chol_batch", "chol_streams", "chol_pipe", "chol_lag",
    "chol_prio", "cf_debug"); None restores the default."""
    check(lib.gpbt_set_option(key.encode(), None if value is None else str(value).encode()))


class on_device:
    """`with on_device(i):` makes CUDA device i current for the enclosed C-ABI calls"""

    def __init__(self, device):
        self.device = device

    def __enter__(self):
        self.prev = lib.gpbt_get_device()
        if self.device is not None and self.device != self.prev:
            check(lib.gpbt_set_device(int(self.device)))
        return self

    def __exit__(self, *exc):
        if self.device is not None and self.device != self.prev:
            lib.gpbt_set_device(self.prev)


def host_ptr(a):
    """pointer to a C-contiguous numpy array (None -> NULL)"""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
