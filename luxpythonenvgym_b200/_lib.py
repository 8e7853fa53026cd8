"""ctypes binding of the C ABI in include/lux_b200.h (luxpythonenvgym_b200/_lux_b200.so).

There is no CPU fallback: if the shared library is missing this raises, and if no CUDA
device is usable the entry points return LUX_E_NO_DEVICE which `check()` turns into an error.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_lux_b200.so")

LUX_STEP_PREVALIDATE = 1
LUX_OK, LUX_E_ARG, LUX_E_CAPACITY, LUX_E_CUDA, LUX_E_NO_DEVICE = 0, -1, -2, -3, -4
_ERR = {LUX_E_ARG: "bad argument", LUX_E_CAPACITY: "capacity out of range",
        LUX_E_CUDA: "CUDA error", LUX_E_NO_DEVICE: "no CUDA device"}

# every symbol include/lux_b200.h declares
EXPORTS = ("lux_b200_version", "lux_b200_device_count", "lux_b200_set_option", "lux_b200_step_smem_bytes",
           "lux_b200_step", "lux_b200_step_host", "lux_b200_gen_actions", "lux_b200_observe",
           "lux_b200_reset_done", "lux_b200_stats", "lux_b200_encode_actions", "lux_b200_reward", "lux_b200_launch_count", "lux_b200_step_host_sparse", "lux_b200_step_host_sparse_submit", "lux_b200_step_host_wait",
           "lux_b200_debug_trace", "lux_b200_reset_gen", "lux_b200_reset_done_dev", "lux_b200_observe_offsets", "lux_b200_env_finish")


class LuxError(RuntimeError):
    def __init__(self, code, what):
        super().__init__("%s: %s (%d)" % (what, _ERR.get(code, "error"), code))
        self.code = code


class LuxBatch(ctypes.Structure):
    _fields_ = [("n_matches", ctypes.c_int32), ("size", ctypes.c_int32),
                ("u_cap", ctypes.c_int32), ("c_cap", ctypes.c_int32),
                ("t_cap", ctypes.c_int32), ("r_cap", ctypes.c_int32),
                ("hdr", ctypes.c_void_p), ("cells", ctypes.c_void_p), ("units", ctypes.c_void_p),
                ("cities", ctypes.c_void_p), ("tiles", ctypes.c_void_p), ("res", ctypes.c_void_p)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "%s is missing: build it with `python -m luxpythonenvgym_b200.build` "
            "(nvcc, sm_100a); there is no CPU fallback" % SO_PATH)
    lib = ctypes.CDLL(SO_PATH)
    vp, i32, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
    bp = ctypes.POINTER(LuxBatch)
    lib.lux_b200_version.restype = i32
    lib.lux_b200_launch_count.restype = ctypes.c_uint64
    lib.lux_b200_device_count.restype = i32
    lib.lux_b200_set_option.argtypes = [ctypes.c_char_p, i32]
    lib.lux_b200_step_smem_bytes.restype = ctypes.c_size_t
    lib.lux_b200_step_smem_bytes.argtypes = [i32] * 5
    lib.lux_b200_step.argtypes = [bp, vp, vp, ctypes.c_uint32, vp]
    lib.lux_b200_step_host.argtypes = [bp, vp, vp, vp, vp, vp, vp, i32, i32, ctypes.c_uint32, vp]
    lib.lux_b200_step_host_sparse.argtypes = [bp, vp, i32, vp, vp, vp, vp, ctypes.c_uint32, vp]
    lib.lux_b200_step_host_sparse_submit.argtypes = [bp, vp, i32, vp, vp, ctypes.c_uint32, vp]
    lib.lux_b200_step_host_wait.argtypes = [bp, vp, vp, vp]
    lib.lux_b200_gen_actions.argtypes = [bp, u64, u64, vp, vp, vp, vp]
    lib.lux_b200_observe.argtypes = [bp, vp, vp, vp, vp, i32, vp, vp]
    lib.lux_b200_reset_done.argtypes = [bp, bp, ctypes.c_uint32, vp, vp]
    lib.lux_b200_reset_done_dev.argtypes = [bp, bp, vp, vp, vp]
    lib.lux_b200_observe_offsets.argtypes = [bp, vp, ctypes.c_int64, vp, vp, vp]
    lib.lux_b200_env_finish.argtypes = [bp, bp, vp, vp, vp, vp, vp, vp, vp]
    lib.lux_b200_reset_gen.argtypes = [bp, bp, ctypes.c_uint32, vp, u64, u64, vp, vp, vp, vp]
    lib.lux_b200_stats.argtypes = [bp, vp, vp]
    lib.lux_b200_encode_actions.argtypes = [bp, vp, vp, vp, i32, vp, vp, vp, vp]
    lib.lux_b200_reward.argtypes = [bp, vp, vp, vp, vp]
    _lib = lib
    return lib


def check(code, what):
    if code != LUX_OK:
        raise LuxError(code, what)
