"""ctypes binding of libpcbgpu.so (include/pcb_env.h).  No fallback: a missing
library or symbol raises ImportError/RuntimeError with the build command."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCB_LIB_PATH") or os.path.join(_HERE, "libpcbgpu.so")
ABI_VERSION = 1

SYMBOLS = (
    "pcb_create", "pcb_destroy", "pcb_upload_instances", "pcb_generate_instances", "pcb_reset",
    "pcb_step", "pcb_sample_random_actions", "pcb_step_host", "pcb_get_tensor",
    "pcb_current_slice", "pcb_episode_stats", "pcb_get_occupancy", "pcb_launch_count",
    "pcb_last_error", "pcb_abi_version", "pcb_rollout_random", "pcb_enable_kernel_timing",
    "pcb_kernel_times", "pcb_current_result_half", "pcb_generate_host", "pcb_generate_instances_device", "pcb_state_layout", "pcb_get_state",
    "pcb_set_state", "pcb_set_reward_overlap", "pcb_shape_id", "pcb_set_fused_rollout",
    "pcb_fused_rollout_times", "pcb_rollout_random_logged", "pcb_check_guards", "pcb_steps_host",
    "pcb_set_graph_rollout", "pcb_graph_replays", "pcb_step_host_packed",
)


class PcbConfig(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("kind", ctypes.c_int32),
        ("height", ctypes.c_int32), ("width", ctypes.c_int32),
        ("component_n", ctypes.c_int32), ("max_num_components", ctypes.c_int32),
        ("max_component_h", ctypes.c_int32), ("max_component_w", ctypes.c_int32),
        ("max_num_nets", ctypes.c_int32), ("max_pins", ctypes.c_int32),
        ("reward_type", ctypes.c_int32), ("reward_beam_width", ctypes.c_int32),
        ("norm_uses_fma", ctypes.c_int32), ("autoreset", ctypes.c_int32),
        ("obs_ring", ctypes.c_int32), ("reserved0", ctypes.c_int32),
        ("weight_wirelength", ctypes.c_double), ("weight_num_intersections", ctypes.c_double),
        ("wirelength_norm", ctypes.c_double), ("intersections_norm", ctypes.c_double),
        ("penalty_reward", ctypes.c_double), ("penalty_wirelength", ctypes.c_double),
        ("penalty_intersections", ctypes.c_double),
    ]


class PcbInstanceBank(ctypes.Structure):
    _fields_ = [
        ("count", ctypes.c_int64),
        ("n_comp", ctypes.c_void_p), ("comp_h", ctypes.c_void_p), ("comp_w", ctypes.c_void_p),
        ("n_pins", ctypes.c_void_p), ("n_nets", ctypes.c_void_p),
        ("pin_comp", ctypes.c_void_p), ("pin_rx", ctypes.c_void_p), ("pin_ry", ctypes.c_void_p),
        ("pin_net", ctypes.c_void_p), ("pin_id", ctypes.c_void_p),
    ]


_libs = {}


def load(path=None):
    """Loads libpcbgpu.so (or a specialised build of it) once; raises with the build hint
    when it is missing."""
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  pcbgpu has no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name in SYMBOLS:
        if not hasattr(lib, name):
            raise ImportError(f"{path} does not export {name}; rebuild the library")
    vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64
    lib.pcb_create.argtypes = [ctypes.POINTER(PcbConfig), i32, i32, ctypes.POINTER(vp)]
    lib.pcb_destroy.argtypes = [vp]
    lib.pcb_upload_instances.argtypes = [vp, ctypes.POINTER(PcbInstanceBank)]
    lib.pcb_generate_instances.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), u64, i64]
    lib.pcb_generate_instances_device.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), u64, i64, vp]
    lib.pcb_reset.argtypes = [vp, vp, vp]
    lib.pcb_step.argtypes = [vp, vp, i32, vp]
    lib.pcb_sample_random_actions.argtypes = [vp, u64, vp, vp]
    lib.pcb_step_host.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.pcb_steps_host.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    lib.pcb_step_host_packed.argtypes = [vp, vp, i32, vp, vp]
    lib.pcb_get_tensor.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp)]
    lib.pcb_current_slice.argtypes = [vp]
    lib.pcb_current_result_half.argtypes = [vp]
    lib.pcb_state_layout.argtypes = [vp, ctypes.POINTER(ctypes.c_int32)]
    lib.pcb_get_state.argtypes = [vp, vp, ctypes.POINTER(u64), vp]
    lib.pcb_set_state.argtypes = [vp, vp, u64, vp]
    lib.pcb_generate_host.argtypes = [i32, i32, i32, i32, ctypes.POINTER(ctypes.c_int32), u64, i64] + [vp] * 10
    lib.pcb_episode_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_double), i32, vp]
    lib.pcb_get_occupancy.argtypes = [vp, vp, vp]
    lib.pcb_launch_count.argtypes = [vp]
    lib.pcb_launch_count.restype = i64
    lib.pcb_rollout_random.argtypes = [vp, u64, i32, vp, vp]
    lib.pcb_check_guards.argtypes = [vp, vp]
    lib.pcb_check_guards.restype = i64
    lib.pcb_rollout_random_logged.argtypes = [vp, u64, i32, vp, vp, vp, vp]
    lib.pcb_enable_kernel_timing.argtypes = [vp, i32]
    lib.pcb_set_reward_overlap.argtypes = [vp, i32]
    lib.pcb_set_fused_rollout.argtypes = [vp, i32]
    lib.pcb_set_graph_rollout.argtypes = [vp, i32]
    lib.pcb_graph_replays.argtypes = [vp]
    lib.pcb_graph_replays.restype = i64
    lib.pcb_fused_rollout_times.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64),
                                            ctypes.POINTER(i64)]
    lib.pcb_kernel_times.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64)]
    lib.pcb_last_error.restype = ctypes.c_char_p
    lib.pcb_abi_version.restype = i32
    lib.pcb_shape_id.argtypes = [vp]
    if lib.pcb_abi_version() != ABI_VERSION:
        raise ImportError(f"libpcbgpu ABI {lib.pcb_abi_version()} != {ABI_VERSION}; rebuild")
    _libs[path] = lib
    return lib


def specialized_library(spec, build=True):
    """Path of a libpcbgpu build whose kernels have every dimension of ``spec`` as a
    compile-time constant (-DPCB_US_KIND=.. -DPCB_US_H=.. ...); compiled with nvcc on first use
    (about a minute) and cached next to the default library."""
    import subprocess

    from .config import KINDS

    c = spec.c
    K = c.max_num_nets + 1 if KINDS[spec.kind] >= 2 else 0
    dims = (KINDS[spec.kind], c.height, c.width, c.max_num_components, c.max_component_h,
            c.max_component_w, K, c.max_pins if KINDS[spec.kind] >= 2 else 0)
    path = os.path.join(_HERE, "libpcbgpu_shape_" + "_".join(str(d) for d in dims) + ".so")
    src = [os.path.join(_HERE, "csrc", n) for n in ("pcb_api.cu", "pcb_gen.cpp")]
    newest = max(os.path.getmtime(os.path.join(_HERE, "csrc", n)) for n in os.listdir(
        os.path.join(_HERE, "csrc")))
    if not os.path.exists(path) or os.path.getmtime(path) < newest:
        if not build:
            raise ImportError(f"{path} is not built")
        cmd = [os.environ.get("NVCC", "nvcc"), "-gencode", "arch=compute_100a,code=sm_100a",
               "-lineinfo", "-O3", "-fmad=false", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
               ] + [f"-DPCB_US_{n}={d}" for n, d in zip(
                   ("KIND", "H", "W", "C", "MH", "MW", "K", "PINS"), dims)] + ["-o", path] + src
        subprocess.check_call(cmd)
    return path


class PcbError(RuntimeError):
    pass


def check(rc, lib=None):
    if rc < 0:
        msg = (lib or load()).pcb_last_error()
        raise PcbError(f"libpcbgpu error {rc}: {msg.decode() if msg else '?'}")
    return rc


_PyCapsule_New = ctypes.pythonapi.PyCapsule_New
_PyCapsule_New.restype = ctypes.py_object
_PyCapsule_New.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]


def tensor(handle, key, lib=None):
    """pcb_get_tensor -> torch tensor sharing the library's device memory."""
    import torch.utils.dlpack

    lib = lib or load()
    out = ctypes.c_void_p()
    check(lib.pcb_get_tensor(handle, key.encode(), ctypes.byref(out)), lib)
    return torch.utils.dlpack.from_dlpack(_PyCapsule_New(out, b"dltensor", None))
