"""ctypes binding of ``libfreddi_b200.so`` (the C ABI of include/freddi_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``freddi_b200.build.build()``.  There is no
fallback: if it is missing, importing this module raises, and without a CUDA device every compute call
returns ``FB200_ERR_NO_DEVICE``.
"""
import ctypes as C
import os

from .args import FB200Args, FB200Config, FB200ModelInfo

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('FB200_LIB') or os.path.join(HERE, 'libfreddi_b200.so')  # FB200_LIB: kernel-variant experiments

FB200_OK = 0
ERR_INVALID_ARGUMENT, ERR_RUNTIME, ERR_LOGIC, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_BAD_HANDLE = -1, -2, -3, -4, -5, -6, -7
MODEL_OK, MODEL_COLLAPSED, MODEL_FAILED = 0, 1, 2

# every symbol include/freddi_b200.h declares
SYMBOLS = (
    'fb200_abi_version', 'fb200_last_error', 'fb200_device_count', 'fb200_measure_fp64_peak', 'fb200_lx_table_selftest', 'fb200_ensemble_counters', 'fb200_args_default', 'fb200_config_default',
    'fb200_args_resolve', 'fb200_ensemble_create', 'fb200_ensemble_destroy', 'fb200_ensemble_n_models',
    'fb200_ensemble_n_cols', 'fb200_ensemble_n_rows', 'fb200_ensemble_nx', 'fb200_ensemble_info', 'fb200_ensemble_evolve',
    'fb200_ensemble_evolve_async', 'fb200_ensemble_reset', 'fb200_ensemble_sync', 'fb200_ensemble_last_evolve_ms', 'fb200_ensemble_last_launches',
    'fb200_ensemble_rows', 'fb200_ensemble_rows_device', 'fb200_ensemble_status', 'fb200_ensemble_state',
    'fb200_ensemble_row_bounds', 'fb200_ensemble_field', 'fb200_ensemble_flux', 'fb200_ensemble_mdot_wind',
    'fb200_ensemble_field_rows', 'fb200_ensemble_flux_rows', 'fb200_ensemble_work_counters', 'fb200_selftest_trapz',
    'fb200_ensemble_device_ptr', 'fb200_ensemble_checkpoint_bytes', 'fb200_ensemble_checkpoint_save', 'fb200_ensemble_checkpoint_load',
    'fb200_ensemble_get_state', 'fb200_ensemble_set_state',
    'fb200_sharded_create', 'fb200_sharded_destroy', 'fb200_sharded_n_shards', 'fb200_sharded_shard', 'fb200_sharded_n_models',
    'fb200_sharded_n_rows', 'fb200_sharded_n_cols', 'fb200_sharded_nx', 'fb200_sharded_info', 'fb200_sharded_reset',
    'fb200_sharded_evolve', 'fb200_sharded_evolve_async', 'fb200_sharded_sync', 'fb200_sharded_last_evolve_ms',
    'fb200_sharded_rows', 'fb200_sharded_status', 'fb200_sharded_state', 'fb200_sharded_work_counters', 'fb200_sharded_field',
    'fb200_sharded_flux', 'fb200_sweep_plan', 'fb200_sweep_run', 'fb200_host_alloc', 'fb200_host_free', 'fb200_trim',
)


class FB200SweepStats(C.Structure):
    """fb200_sweep_stats_t (include/freddi_b200.h)."""
    _fields_ = [('total_ms', C.c_double), ('device_ms_max', C.c_double), ('first_launch_ms', C.c_double), ('resolve_ms', C.c_double),
                ('uploads_ms', C.c_double), ('model_steps', C.c_int64), ('n_devices', C.c_int32), ('launches', C.c_int32),
                ('n_rows', C.c_int32), ('n_cols', C.c_int32)]


class FB200Error(Exception):
    """Base of the errors raised from a failing C-ABI call; ``code`` is the FB200_ERR_* value."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class NoDeviceError(FB200Error, RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            '{} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(nvcc, sm_100a). freddi_b200 has no CPU fallback.'.format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    dp, ip, i32p = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    vp = C.c_void_p
    lib.fb200_abi_version.restype = C.c_int
    lib.fb200_last_error.restype = C.c_char_p
    lib.fb200_device_count.restype = C.c_int
    lib.fb200_lx_table_selftest.argtypes = [C.c_double] * 4 + [C.c_long, C.c_ulonglong, dp, C.POINTER(C.c_long), dp, C.POINTER(C.c_int)]
    lib.fb200_args_default.argtypes = [C.POINTER(FB200Args), C.c_int]
    lib.fb200_args_default.restype = None
    lib.fb200_config_default.argtypes = [C.POINTER(FB200Config)]
    lib.fb200_config_default.restype = None
    lib.fb200_args_resolve.argtypes = [C.POINTER(FB200Args), C.POINTER(FB200ModelInfo), dp, dp]
    lib.fb200_ensemble_create.argtypes = [C.POINTER(FB200Args), C.c_size_t, C.POINTER(FB200Config), C.POINTER(vp),
                                          C.POINTER(C.c_size_t)]
    lib.fb200_ensemble_destroy.argtypes = [vp]
    lib.fb200_ensemble_destroy.restype = None
    for name in ('n_models', 'n_cols', 'n_rows', 'nx'):
        f = getattr(lib, 'fb200_ensemble_' + name)
        f.argtypes = [vp]
        f.restype = C.c_size_t
    lib.fb200_ensemble_info.argtypes = [vp, C.POINTER(FB200ModelInfo)]
    lib.fb200_ensemble_evolve.argtypes = [vp, C.c_int64]
    lib.fb200_ensemble_evolve_async.argtypes = [vp, C.c_int64]
    lib.fb200_ensemble_sync.argtypes = [vp]
    lib.fb200_ensemble_reset.argtypes = [vp]
    lib.fb200_ensemble_last_evolve_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.fb200_ensemble_last_launches.argtypes = [vp]
    lib.fb200_ensemble_rows.argtypes = [vp, dp, C.c_size_t]
    lib.fb200_ensemble_rows_device.argtypes = [vp, C.POINTER(vp)]
    lib.fb200_ensemble_status.argtypes = [vp, i32p, i32p, ip]
    lib.fb200_ensemble_state.argtypes = [vp, dp, ip, ip, ip]
    lib.fb200_ensemble_counters.argtypes = [vp, ip, ip]
    lib.fb200_ensemble_work_counters.argtypes = [vp, ip, C.c_size_t]
    lib.fb200_measure_fp64_peak.argtypes = [C.c_int, dp]
    lib.fb200_selftest_trapz.argtypes = [C.c_int, dp, dp, C.c_size_t, C.c_size_t, C.c_size_t, dp]
    lib.fb200_ensemble_row_bounds.argtypes = [vp, C.c_int64, ip, ip]
    lib.fb200_ensemble_field.argtypes = [vp, C.c_int, C.c_int64, dp, C.c_size_t]
    lib.fb200_ensemble_flux.argtypes = [vp, C.c_int, C.c_int64, dp, C.c_size_t, dp, C.c_size_t]
    lib.fb200_ensemble_mdot_wind.argtypes = [vp, C.c_int64, dp, C.c_size_t]
    lib.fb200_ensemble_field_rows.argtypes = [vp, C.c_int, C.c_int64, C.c_int64, C.c_int, dp, C.c_size_t]
    lib.fb200_ensemble_flux_rows.argtypes = [vp, C.c_int, C.c_int64, C.c_int64, dp, C.c_size_t, dp, dp, C.c_size_t]
    lib.fb200_ensemble_device_ptr.argtypes = [vp, C.c_int, C.POINTER(vp), i32p]
    lib.fb200_ensemble_checkpoint_bytes.argtypes = [vp, C.c_int]
    lib.fb200_ensemble_checkpoint_bytes.restype = C.c_size_t
    lib.fb200_ensemble_checkpoint_save.argtypes = [vp, vp, C.c_size_t, C.c_int]
    lib.fb200_ensemble_checkpoint_load.argtypes = [vp, vp, C.c_size_t]
    lib.fb200_ensemble_get_state.argtypes = [vp, dp, dp, ip, ip, ip, dp, dp]
    lib.fb200_ensemble_set_state.argtypes = [vp, dp, dp, ip, ip, ip, dp, dp]
    # in-process multi-GPU + one-shot sweep
    szp = C.POINTER(C.c_size_t)
    lib.fb200_sharded_create.argtypes = [C.POINTER(FB200Args), C.c_size_t, C.POINTER(FB200Config), i32p, C.c_int32, C.POINTER(vp), szp]
    lib.fb200_sharded_destroy.argtypes = [vp]
    lib.fb200_sharded_destroy.restype = None
    lib.fb200_sharded_n_shards.argtypes = [vp]
    lib.fb200_sharded_n_shards.restype = C.c_int32
    lib.fb200_sharded_shard.argtypes = [vp, C.c_int32]
    lib.fb200_sharded_shard.restype = vp
    for name in ('n_models', 'n_cols', 'n_rows', 'nx'):
        f = getattr(lib, 'fb200_sharded_' + name)
        f.argtypes = [vp]
        f.restype = C.c_size_t
    lib.fb200_sharded_info.argtypes = [vp, C.POINTER(FB200ModelInfo)]
    lib.fb200_sharded_reset.argtypes = [vp]
    lib.fb200_sharded_evolve.argtypes = [vp, C.c_int64]
    lib.fb200_sharded_evolve_async.argtypes = [vp, C.c_int64]
    lib.fb200_sharded_sync.argtypes = [vp]
    lib.fb200_sharded_last_evolve_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.fb200_sharded_rows.argtypes = [vp, dp, C.c_size_t]
    lib.fb200_sharded_status.argtypes = [vp, i32p, i32p, ip]
    lib.fb200_sharded_state.argtypes = [vp, dp, ip, ip, ip]
    lib.fb200_sharded_work_counters.argtypes = [vp, ip, C.c_size_t]
    lib.fb200_sharded_field.argtypes = [vp, C.c_int, C.c_int64, dp, C.c_size_t]
    lib.fb200_sharded_flux.argtypes = [vp, C.c_int, C.c_int64, dp, C.c_size_t, dp, C.c_size_t]
    lib.fb200_sweep_plan.argtypes = [C.POINTER(FB200Args), C.c_size_t, C.POINTER(FB200Config), szp, szp]
    lib.fb200_sweep_run.argtypes = [C.POINTER(FB200Args), C.c_size_t, C.POINTER(FB200Config), i32p, C.c_int32, dp, C.c_size_t,
                                    i32p, i32p, ip, dp, C.POINTER(FB200SweepStats), szp]
    lib.fb200_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    lib.fb200_host_free.argtypes = [vp]
    lib.fb200_host_free.restype = None
    lib.fb200_trim.argtypes = []
    lib.fb200_trim.restype = None
    return lib


lib = _load()


def check(rc):
    """Re-raise a failing call as the exception class the reference would have thrown
    (std::invalid_argument -> ValueError, std::runtime_error -> RuntimeError, std::logic_error -> RuntimeError)."""
    if rc == FB200_OK:
        return
    msg = (lib.fb200_last_error() or b'').decode()
    if rc == ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    if rc == ERR_NO_DEVICE:
        raise NoDeviceError(rc, msg)
    if rc in (ERR_RUNTIME, ERR_LOGIC):
        raise RuntimeError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise FB200Error(rc, msg)
