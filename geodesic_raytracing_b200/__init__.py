"""geodesic_raytracing_b200 - host-side Python mirror of the reference's plugin + frame interface.

Everything here is a thin ctypes layer over ``libgrt.so`` (the C ABI declared in ``include/grt.h``):

* :class:`Content`  ~ ``content_manager`` (/root/reference/content_manager.hpp:47-58)
* :class:`Metric`   ~ ``metrics::metric``  (/root/reference/metric.hpp:330-435, 624-663)
* :class:`Context`  ~ the frame loop + ``render_state`` (/root/reference/main.cpp:2244-2526,
  /root/reference/render_state.hpp:94-197)

There is no CPU implementation behind these classes: device calls raise :class:`GrtError` when the
CUDA library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgrt.so")
CONTENT_DIR = os.path.join(_HERE, "content")

GRT_FORMAT_RGBA32F = 0
GRT_FORMAT_RGBA8 = 1
GRT_COMM_HANDLE_BYTES = 64

COORDINATE_SYSTEMS = ("X_Y_THETA_PHI", "CARTESIAN", "CYLINDRICAL", "OTHER")


class GrtError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


class MetricInfo(C.Structure):
    _fields_ = [
        ("is_big", C.c_int), ("is_constant_theta", C.c_int), ("use_prepass", C.c_int), ("singular", C.c_int),
        ("traversable_event_horizon", C.c_int), ("adaptive_precision", C.c_int), ("detect_singularities", C.c_int),
        ("coordinate_system", C.c_int), ("has_periodicity", C.c_int), ("unconditionally_nonsingular", C.c_int),
        ("max_acceleration_change", C.c_float), ("singular_terminator", C.c_float), ("w_v", C.c_float * 4),
        ("flops_expr", C.c_int), ("flops_accel", C.c_int), ("flops_accel_reference", C.c_int),
    ]


class Features(C.Structure):
    """dynamic_feature_config (/root/reference/main.cpp:1123-1158)."""
    _fields_ = [
        ("universe_size", C.c_float), ("max_acceleration_change", C.c_float), ("max_precision_radius", C.c_float),
        ("min_step", C.c_float), ("adaptive_sampling_threshold", C.c_float), ("field_of_view", C.c_float),
        ("redshift", C.c_int), ("reparameterisation", C.c_int), ("use_old_redshift", C.c_int),
        ("adaptive_sampling", C.c_int), ("max_probes", C.c_int), ("use_prepass", C.c_int),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("trial_steps", C.c_uint64), ("accepted_steps", C.c_uint64), ("rays_integrated", C.c_uint64),
        ("rays_finished", C.c_uint64), ("extra_rays", C.c_uint64),
        ("ms_camera", C.c_float), ("ms_prepass", C.c_float), ("ms_init", C.c_float), ("ms_integrate", C.c_float),
        ("ms_render_data", C.c_float), ("ms_adaptive", C.c_float), ("ms_render", C.c_float), ("ms_total", C.c_float),
        ("integrate_launches", C.c_int), ("total_launches", C.c_int),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class CameraState(C.Structure):
    _fields_ = [("pos_generic", C.c_float * 4), ("tetrad", C.c_float * 16), ("timelike_coordinate", C.c_int)]


LIGHTRAY_DTYPE = np.dtype([
    ("position", "<f4", 4), ("velocity", "<f4", 4), ("initial_quat", "<f4", 4), ("acceleration", "<f4", 4),
    ("ku_uobsu", "<f4"), ("running_dlambda_dnew", "<f4"), ("terminated", "<i4"), ("sx", "<i4"), ("sy", "<i4"),
    ("pad", "<i4", 3),
])
RENDER_DATA_DTYPE = np.dtype([
    ("tex_coord", "<f4", 2), ("z_shift", "<f4"), ("sx", "<i4"), ("sy", "<i4"), ("terminated", "<i4"), ("side", "<i4"),
    ("pad", "<i4"),
])
assert LIGHTRAY_DTYPE.itemsize == 96 and RENDER_DATA_DTYPE.itemsize == 32

_lib = None


def build_library(force: bool = False) -> str:
    """Compile libgrt.so in-tree (nvcc, sm_100a). Returns its path."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "-j", str(os.cpu_count() or 4)], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib() -> C.CDLL:
    """Load libgrt.so (fails loudly when it has not been built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GrtError(-4, f"{LIB_PATH} is missing: run `make -C {_HERE}` (or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, cp, ci, cf = C.c_void_p, C.c_char_p, C.c_int, C.c_float
    pf, pi = C.POINTER(C.c_float), C.POINTER(C.c_int)
    sig = {
        "grt_last_error": (cp, []), "grt_version": (cp, []),
        "grt_content_create": (vp, []), "grt_content_destroy": (None, [vp]), "grt_content_add_folder": (ci, [vp, cp]),
        "grt_content_metric_count": (ci, [vp]), "grt_content_metric_name": (cp, [vp, ci]),
        "grt_metric_load": (vp, [vp, cp]), "grt_metric_load_by_name": (vp, [vp, cp]), "grt_metric_destroy": (None, [vp]),
        "grt_metric_param_count": (ci, [vp]), "grt_metric_param_name": (cp, [vp, ci]), "grt_metric_param_default": (cf, [vp, ci]),
        "grt_metric_param_get": (cf, [vp, ci]), "grt_metric_param_set": (ci, [vp, ci, cf]),
        "grt_metric_get_info": (ci, [vp, C.POINTER(MetricInfo)]), "grt_metric_name": (cp, [vp]), "grt_metric_description": (cp, [vp]),
        "grt_image_load_png": (vp, [cp, C.POINTER(ci), C.POINTER(ci)]), "grt_image_free": (None, [vp]), "grt_image_save_png": (ci, [cp, vp, ci, ci]),
        "grt_metric_packed_params": (ci, [vp, vp, ci]),
        "grt_metric_source": (cp, [vp, ci]), "grt_metric_eval_symbolic": (ci, [vp] + [vp] * 9), "grt_compile_metric": (ci, [vp, ci, ci, C.POINTER(cp)]),
        "grt_features_default": (None, [C.POINTER(Features)]),
        "grt_create": (vp, [ci]), "grt_destroy": (None, [vp]), "grt_set_metric": (ci, [vp, vp, ci]), "grt_update_params": (ci, [vp]), "grt_set_metric_async": (ci, [vp, vp, ci]), "grt_build_poll": (ci, [vp]),
        "grt_set_fast_math": (ci, [vp, ci]), "grt_set_tuning": (ci, [vp, ci, ci]), "grt_set_frames_in_flight": (ci, [vp, ci]), "grt_set_features": (ci, [vp, C.POINTER(Features)]), "grt_get_features": (ci, [vp, C.POINTER(Features)]),
        "grt_set_background": (ci, [vp, vp, ci, ci, ci]), "grt_set_background_device": (ci, [vp, vp, ci, ci, ci]),
        "grt_set_camera": (ci, [vp, pf, pf, ci]), "grt_set_observer_speed": (ci, [vp, pf]),
        "grt_snapshot_geodesic": (ci, [vp, cf, ci]), "grt_set_camera_on_geodesic": (ci, [vp, cf]), "grt_clear_camera_geodesic": (ci, [vp]),
        "grt_render": (ci, [vp, ci, ci, vp, ci, ci, C.POINTER(Stats)]),
        "grt_render_rows": (ci, [vp, ci, ci, ci, ci, vp, ci, ci, C.POINTER(Stats)]),
        "grt_render_stripes": (ci, [vp, ci, ci, pi, pi, ci, vp, ci, ci, C.POINTER(Stats)]),
        "grt_synchronize": (ci, [vp]), "grt_stream": (vp, [vp]),
        "grt_comm_init": (ci, [vp, ci, ci, ci, ci, ci, ci, vp]), "grt_comm_connect": (ci, [vp, vp]), "grt_comm_connect_local": (ci, [C.POINTER(vp), ci]), "grt_comm_disconnect": (ci, [vp]),
        "grt_comm_render": (ci, [vp, pi, pi, ci, C.c_uint, vp]), "grt_comm_frame": (vp, [vp, C.c_uint]),
        "grt_interleaved_stripes": (ci, [ci, ci, ci, ci, pi, pi, ci]),
        "grt_measure_stripe_costs": (ci, [vp, ci, ci, ci, C.POINTER(C.c_double), ci]),
        "grt_balanced_stripes": (ci, [ci, ci, ci, ci, C.POINTER(C.c_double), ci, pi, pi, ci]),
        "grt_contiguous_bounds": (ci, [ci, ci, ci, C.POINTER(C.c_double), ci, pi]),
        "grt_refine_bounds": (ci, [ci, ci, ci, C.POINTER(C.c_double), ci, pi, C.POINTER(C.c_double), pi]),
        "grt_group_create": (vp, [pi, ci]), "grt_group_destroy": (None, [vp]), "grt_group_size": (ci, [vp]), "grt_group_ctx": (vp, [vp, ci]),
        "grt_group_set_metric": (ci, [vp, vp, ci]), "grt_group_update_params": (ci, [vp]), "grt_group_set_features": (ci, [vp, C.POINTER(Features)]),
        "grt_group_set_background": (ci, [vp, vp, ci, ci, ci]), "grt_group_set_camera": (ci, [vp, pf, pf, ci]), "grt_group_set_frames_in_flight": (ci, [vp, ci]),
        "grt_group_render": (ci, [vp, ci, ci, ci, vp]), "grt_group_frame": (vp, [vp]), "grt_group_synchronize": (ci, [vp]),
        "grt_group_rebalance": (ci, [vp, ci, ci]),
        "grt_get_camera_state": (ci, [vp, C.POINTER(CameraState)]),
        "grt_debug_ray_count": (ci, [vp, pi, pi]), "grt_debug_read_rays": (ci, [vp, vp, ci]), "grt_debug_read_render_data": (ci, [vp, vp, ci]),
        "grt_debug_read_prepass_mask": (ci, [vp, vp, ci, pi, pi]), "grt_debug_read_mip_level": (ci, [vp, ci, ci, vp, pi, pi, pi]),
        "grt_debug_read_geodesic": (ci, [vp, vp, vp, vp, vp, ci]),
        "grt_debug_eval_metric": (ci, [vp, vp, vp, ci, vp, vp, vp, vp, vp]), "grt_debug_tetrad_at": (ci, [vp, vp, ci, pf, ci, vp, vp]),
        "grt_debug_integrate_launch_ms": (ci, [vp, pf, ci]),
        "grt_debug_kernel_info": (cp, [vp]), "grt_debug_launch_log": (cp, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily by exported_symbols()


def _err() -> str:
    return (lib().grt_last_error() or b"").decode("utf-8", "replace")


def _check(rc: int) -> int:
    if rc < 0:
        raise GrtError(rc, _err())
    return rc


def _f32(a, n=None) -> np.ndarray:
    arr = np.ascontiguousarray(a, dtype=np.float32)
    if n is not None and arr.size != n:
        raise ValueError(f"expected {n} floats, got {arr.size}")
    return arr


def default_features() -> Features:
    f = Features()
    lib().grt_features_default(C.byref(f))
    return f


class Content:
    """A set of content folders (scripts + configs)."""

    def __init__(self, *folders: str):
        self._h = lib().grt_content_create()
        for f in folders:
            self.add_folder(f)

    def add_folder(self, folder: str) -> None:
        _check(lib().grt_content_add_folder(self._h, os.fsencode(folder)))

    def metric_names(self) -> List[str]:
        L = lib()
        return [L.grt_content_metric_name(self._h, i).decode() for i in range(L.grt_content_metric_count(self._h))]

    def load(self, js_path: str) -> "Metric":
        h = lib().grt_metric_load(self._h, os.fsencode(js_path))
        if not h:
            raise GrtError(-2, _err())
        return Metric(h, self)

    def load_by_name(self, name: str) -> "Metric":
        h = lib().grt_metric_load_by_name(self._h, name.encode())
        if not h:
            raise GrtError(-2, _err())
        return Metric(h, self)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.grt_content_destroy(self._h)
            self._h = None


class Metric:
    def __init__(self, handle, content: Content):
        self._h = handle
        self._content = content  # keep alive

    @property
    def name(self) -> str:
        return lib().grt_metric_name(self._h).decode()

    @property
    def description(self) -> str:
        return lib().grt_metric_description(self._h).decode()

    @property
    def info(self) -> MetricInfo:
        out = MetricInfo()
        _check(lib().grt_metric_get_info(self._h, C.byref(out)))
        return out

    @property
    def param_names(self) -> List[str]:
        L = lib()
        return [L.grt_metric_param_name(self._h, i).decode() for i in range(L.grt_metric_param_count(self._h))]

    @property
    def param_defaults(self) -> List[float]:
        L = lib()
        return [L.grt_metric_param_default(self._h, i) for i in range(L.grt_metric_param_count(self._h))]

    @property
    def params(self) -> List[float]:
        L = lib()
        return [L.grt_metric_param_get(self._h, i) for i in range(L.grt_metric_param_count(self._h))]

    @property
    def packed_params(self) -> List[float]:
        """`struct grt_params` of the generated source: the named values followed by the derived parameters."""
        n = _check(lib().grt_metric_packed_params(self._h, None, 0))
        buf = (C.c_float * max(1, n))()
        _check(lib().grt_metric_packed_params(self._h, buf, n))
        return list(buf[:n])

    def set_param(self, name_or_index, value: float) -> None:
        idx = name_or_index if isinstance(name_or_index, int) else self.param_names.index(name_or_index)
        _check(lib().grt_metric_param_set(self._h, idx, float(value)))

    def source(self, is_static: bool = False) -> str:
        s = lib().grt_metric_source(self._h, int(is_static))
        if s is None:
            raise GrtError(-2, _err())
        return s.decode()

    def eval_symbolic(self, pos, vel=(0, 0, 0, 0)) -> dict:
        """Host-side float64 evaluation of the metric's symbolic expressions (plugin-layer diagnostic, grt.h)."""
        p, v = np.ascontiguousarray(pos, np.float64), np.ascontiguousarray(vel, np.float64)
        out = {"g": np.zeros(16), "dg": np.zeros(64), "accel": np.zeros(4), "accel_ref": np.zeros(4), "to_polar": np.zeros(4), "from_polar": np.zeros(4), "distance": np.zeros(1)}
        _check(lib().grt_metric_eval_symbolic(self._h, p.ctypes.data, v.ctypes.data, *[a.ctypes.data for a in out.values()]))
        return out

    def compile(self, is_static: bool = False, fast_math: bool = True) -> str:
        """NVRTC-compile for sm_100a into the cubin cache (needs no GPU). Returns the build log."""
        log = C.c_char_p()
        _check(lib().grt_compile_metric(self._h, int(is_static), int(fast_math), C.byref(log)))
        return (log.value or b"").decode("utf-8", "replace")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.grt_metric_destroy(self._h)
            self._h = None


def _ptr(x) -> int:
    """Device or host address of a numpy array / torch tensor / raw int."""
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


class Context:
    """One GPU: the bound metric program, device buffers, camera, and the per-frame kernel sequence."""

    def __init__(self, device: int = 0):
        self._h = lib().grt_create(int(device))
        if not self._h:
            raise GrtError(-4, _err())
        self.device = device
        self._metric = None

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.grt_destroy(self._h)
        self._h = None

    __del__ = close

    # ---- program / features -----------------------------------------------------------------
    def set_fast_math(self, enabled: bool) -> None:
        _check(lib().grt_set_fast_math(self._h, int(enabled)))

    def measure_stripe_costs(self, width: int, height: int, stripe_rows: int = 54) -> List[float]:
        """Trial ray-steps of every half stripe of the frame rendered by itself (grt_measure_stripe_costs; two numbers per stripe):
        the cost profile grt_balanced_stripes deals the stripes by."""
        cap = height + 4
        cost = (C.c_double * cap)()
        n = _check(lib().grt_measure_stripe_costs(self._h, width, height, stripe_rows, cost, cap))
        return [cost[i] for i in range(n)]

    def set_tuning(self, blocks_per_sm: int, donate_threshold: int) -> None:
        """Scheduling knobs of the integration kernel (never change results): see grt_set_tuning in include/grt.h."""
        _check(lib().grt_set_tuning(self._h, int(blocks_per_sm), int(donate_threshold)))

    def set_frames_in_flight(self, n: int) -> None:
        _check(lib().grt_set_frames_in_flight(self._h, int(n)))

    def set_metric(self, metric: Metric, is_static: bool = False) -> None:
        _check(lib().grt_set_metric(self._h, metric._h, int(is_static)))
        self._metric = metric

    def set_metric_async(self, metric: Metric, is_static: bool = False) -> None:
        """Build in the background; keep rendering with the bound program and call build_poll() until it returns True."""
        _check(lib().grt_set_metric_async(self._h, metric._h, int(is_static)))
        self._pending_metric = metric

    def build_poll(self) -> bool:
        done = _check(lib().grt_build_poll(self._h)) == 1
        if done and getattr(self, "_pending_metric", None) is not None:
            self._metric, self._pending_metric = self._pending_metric, None
        return done

    def update_params(self) -> None:
        _check(lib().grt_update_params(self._h))

    @property
    def features(self) -> Features:
        f = Features()
        _check(lib().grt_get_features(self._h, C.byref(f)))
        return f

    @features.setter
    def features(self, f: Features) -> None:
        _check(lib().grt_set_features(self._h, C.byref(f)))

    def set_features(self, **kw) -> Features:
        f = self.features
        for k, v in kw.items():
            if not hasattr(f, k):
                raise AttributeError(k)
            setattr(f, k, v)
        self.features = f
        return f

    # ---- inputs -----------------------------------------------------------------------------
    def set_background(self, rgba, which: int = 0, width: Optional[int] = None, height: Optional[int] = None) -> None:
        """rgba: HxWx4 uint8 numpy array (host) or a CUDA torch tensor of the same shape (device)."""
        if isinstance(rgba, np.ndarray):
            a = np.ascontiguousarray(rgba, dtype=np.uint8)
            h, w = a.shape[:2]
            _check(lib().grt_set_background(self._h, a.ctypes.data, w, h, which))
        else:
            h, w = (height, width) if width else rgba.shape[:2]
            _check(lib().grt_set_background_device(self._h, _ptr(rgba), w, h, which))

    def set_camera(self, pos_txyz: Sequence[float] = (0, 0, -4, 0), quat_xyzw: Sequence[float] = (-0.70710678, 0, 0, 0.70710678), flip_r: bool = False) -> None:
        p, q = _f32(pos_txyz, 4), _f32(quat_xyzw, 4)
        _check(lib().grt_set_camera(self._h, p.ctypes.data_as(C.POINTER(C.c_float)), q.ctypes.data_as(C.POINTER(C.c_float)), int(flip_r)))

    def set_observer_speed(self, speed_xyz: Sequence[float]) -> None:
        s = _f32(speed_xyz, 3)
        _check(lib().grt_set_observer_speed(self._h, s.ctypes.data_as(C.POINTER(C.c_float))))

    def snapshot_geodesic(self, linear_basis_speed: float = 0.0, parallel_transport_observer: bool = True) -> int:
        return _check(lib().grt_snapshot_geodesic(self._h, float(linear_basis_speed), int(parallel_transport_observer)))

    def set_camera_on_geodesic(self, proper_time: float) -> None:
        _check(lib().grt_set_camera_on_geodesic(self._h, float(proper_time)))

    def clear_camera_geodesic(self) -> None:
        _check(lib().grt_clear_camera_geodesic(self._h))

    # ---- frames -----------------------------------------------------------------------------
    def render(self, width: int, height: int, fmt: int = GRT_FORMAT_RGBA32F, out=None, stats: bool = True, stripes: Optional[Sequence[Tuple[int, int]]] = None):
        """Render one frame (or the given row stripes). Returns (image, Stats|None).

        out None -> a new host numpy array; a numpy array -> host output (one D2H copy inside the call, synchronous);
        a CUDA torch tensor -> device output, no copy, asynchronous unless stats are requested;
        a pinned CPU torch tensor -> host output, D2H copy queued behind the frame, asynchronous likewise.
        Two frames can be in flight: consecutive asynchronous calls alternate between two streams."""
        if stripes is None:
            stripes = [(0, height)]
        rows = sum(b - a for a, b in stripes)
        dtype, ch = (np.float32, 4) if fmt == GRT_FORMAT_RGBA32F else (np.uint8, 4)
        is_device = 0
        if out is None:
            out = np.empty((rows, width, ch), dtype=dtype)
        elif not isinstance(out, np.ndarray):
            # torch tensor: CUDA -> device output; pinned CPU -> asynchronous host output (call synchronize() before reading)
            is_device = 1 if getattr(out, "is_cuda", True) else 2
        st = Stats() if stats else None
        y0 = (C.c_int * len(stripes))(*[a for a, _ in stripes])
        y1 = (C.c_int * len(stripes))(*[b for _, b in stripes])
        _check(lib().grt_render_stripes(self._h, width, height, y0, y1, len(stripes), _ptr(out), fmt, int(is_device), C.byref(st) if st is not None else None))
        return out, st

    def synchronize(self) -> None:
        _check(lib().grt_synchronize(self._h))

    # ---- one frame over several GPUs (grt.h: grt_comm_*) ------------------------------------------
    def comm_init(self, rank: int, world: int, width: int, height: int, fmt: int = GRT_FORMAT_RGBA8, n_frames: int = 4) -> bytes:
        """Allocate this rank's communication region (the frames too on rank 0); returns its IPC handle (GRT_COMM_HANDLE_BYTES)."""
        h = (C.c_uint8 * GRT_COMM_HANDLE_BYTES)()
        _check(lib().grt_comm_init(self._h, rank, world, width, height, fmt, n_frames, C.addressof(h)))
        return bytes(h)

    def comm_connect(self, handles: bytes) -> None:
        """One process per GPU: `handles` = the handles of all ranks in rank order (world * GRT_COMM_HANDLE_BYTES bytes)."""
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        _check(lib().grt_comm_connect(self._h, C.addressof(buf)))

    @staticmethod
    def comm_connect_local(contexts: Sequence["Context"]) -> None:
        """One process, several contexts (in rank order): peer access instead of IPC."""
        arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
        _check(lib().grt_comm_connect_local(arr, len(contexts)))

    def comm_disconnect(self) -> None:
        """Unmap the other ranks' regions; every rank does this (then a barrier) before any rank closes its context."""
        _check(lib().grt_comm_disconnect(self._h))

    def comm_render(self, stripes: Sequence[Tuple[int, int]], seq: int, host_out=None) -> None:
        """Render this rank's stripes of frame `seq` straight into the owner's frame buffer (asynchronous). On rank 0 `host_out`
        (pinned) receives the assembled frame behind the last rank's stores."""
        y0 = (C.c_int * len(stripes))(*[a for a, _ in stripes])
        y1 = (C.c_int * len(stripes))(*[b for _, b in stripes])
        _check(lib().grt_comm_render(self._h, y0, y1, len(stripes), seq, _ptr(host_out) if host_out is not None else None))

    def integrate_launch_ms(self) -> List[float]:
        """Durations (ms) of the integration launches of the most recent frame of every frame slot in use (waits for them)."""
        buf = (C.c_float * 64)()
        n = _check(lib().grt_debug_integrate_launch_ms(self._h, buf, 64))
        return [buf[i] for i in range(min(n, 64))]

    def comm_frame(self, seq: int) -> int:
        p = lib().grt_comm_frame(self._h, seq)
        if not p:
            raise GrtError(-5, _err())
        return p

    @property
    def stream(self) -> int:
        return lib().grt_stream(self._h) or 0

    # ---- introspection ----------------------------------------------------------------------
    def camera_state(self) -> CameraState:
        s = CameraState()
        _check(lib().grt_get_camera_state(self._h, C.byref(s)))
        return s

    def ray_count(self) -> Tuple[int, int]:
        a, b = C.c_int(), C.c_int()
        _check(lib().grt_debug_ray_count(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def read_rays(self) -> np.ndarray:
        a, b = self.ray_count()
        out = np.zeros(a + b, dtype=LIGHTRAY_DTYPE)
        n = _check(lib().grt_debug_read_rays(self._h, out.ctypes.data, len(out)))
        return out[:n]

    def read_render_data(self, width: int, rows: int) -> np.ndarray:
        out = np.zeros(width * rows, dtype=RENDER_DATA_DTYPE)
        n = _check(lib().grt_debug_read_render_data(self._h, out.ctypes.data, len(out)))
        return out[:n]

    def read_prepass_mask(self) -> Optional[np.ndarray]:
        pw, ph = C.c_int(), C.c_int()
        _check(lib().grt_debug_read_prepass_mask(self._h, None, 0, C.byref(pw), C.byref(ph)))
        if pw.value * ph.value == 0:
            return None
        out = np.zeros(pw.value * ph.value, dtype=np.int32)
        n = _check(lib().grt_debug_read_prepass_mask(self._h, out.ctypes.data, len(out), C.byref(pw), C.byref(ph)))
        return out.reshape(ph.value, pw.value) if n else None

    def read_mip_level(self, which: int, level: int) -> Tuple[np.ndarray, int]:
        lw, lh, levels = C.c_int(), C.c_int(), C.c_int()
        _check(lib().grt_debug_read_mip_level(self._h, which, level, None, C.byref(lw), C.byref(lh), C.byref(levels)))
        out = np.zeros((lh.value, lw.value, 4), dtype=np.uint8)
        _check(lib().grt_debug_read_mip_level(self._h, which, level, out.ctypes.data, C.byref(lw), C.byref(lh), C.byref(levels)))
        return out, levels.value

    def read_geodesic(self, max_points: int = 64000):
        pos = np.zeros((max_points, 4), np.float32)
        vel = np.zeros((max_points, 4), np.float32)
        ds = np.zeros(max_points, np.float32)
        tr = np.zeros((4, max_points, 4), np.float32)
        n = _check(lib().grt_debug_read_geodesic(self._h, pos.ctypes.data, vel.ctypes.data, ds.ctypes.data, tr.ctypes.data, max_points))
        return pos[:n], vel[:n], ds[:n], tr[:, :n]

    def eval_metric(self, positions, velocities):
        p = _f32(positions).reshape(-1, 4)
        v = _f32(velocities).reshape(-1, 4)
        n = len(p)
        g = np.zeros((n, 16), np.float32)
        dg = np.zeros((n, 64), np.float32)
        acc = np.zeros((n, 4), np.float32)
        pol = np.zeros((n, 4), np.float32)
        rad = np.zeros((n, 2), np.float32)
        _check(lib().grt_debug_eval_metric(self._h, p.ctypes.data, v.ctypes.data, n, g.ctypes.data, dg.ctypes.data, acc.ctypes.data, pol.ctypes.data, rad.ctypes.data))
        return g, dg, acc, pol, rad

    def tetrad_at(self, positions, speed=(0, 0, 0), orient: bool = False):
        p = _f32(positions).reshape(-1, 4)
        s = _f32(speed, 3)
        n = len(p)
        tet = np.zeros((n, 4, 4), np.float32)
        tl = np.zeros(n, np.int32)
        _check(lib().grt_debug_tetrad_at(self._h, p.ctypes.data, n, s.ctypes.data_as(C.POINTER(C.c_float)), int(orient), tet.ctypes.data, tl.ctypes.data))
        return tet, tl

    def kernel_info(self) -> dict:
        s = (lib().grt_debug_kernel_info(self._h) or b"").decode()
        out = {}
        for item in filter(None, s.split(";")):
            parts = item.split()
            out[parts[0]] = tuple(int(x) for x in parts[1:])
        return out

    def launch_log(self) -> dict:
        s = (lib().grt_debug_launch_log(self._h) or b"").decode()
        return {k: float(v) for k, v in (item.split(":") for item in filter(None, s.split(";")))}


class Group:
    """Several devices of one box driven as one renderer (grt_group_*): rank 0 = devices[0] owns the assembled frame."""

    def __init__(self, devices: Sequence[int]):
        arr = (C.c_int * len(devices))(*devices)
        self._g = lib().grt_group_create(arr, len(devices))
        if not self._g:
            raise GrtError(-3, _err())
        self.size = len(devices)

    def close(self):
        if getattr(self, "_g", None):
            lib().grt_group_destroy(self._g)
            self._g = None

    __del__ = close

    def set_metric(self, metric: "Metric", is_static: bool = False):
        _check(lib().grt_group_set_metric(self._g, metric._h, int(is_static)))

    def set_features(self, **kw):
        f = Features()
        lib().grt_features_default(C.byref(f))
        for k, v in kw.items():
            if not hasattr(f, k):
                raise AttributeError(k)
            setattr(f, k, v)
        _check(lib().grt_group_set_features(self._g, C.byref(f)))

    def set_background(self, rgba: np.ndarray, which: int = 0):
        a = np.ascontiguousarray(rgba, np.uint8)
        _check(lib().grt_group_set_background(self._g, a.ctypes.data, a.shape[1], a.shape[0], which))

    def set_camera(self, pos=(0.0, 0.0, -4.0, 0.0), quat=(-0.70710678, 0.0, 0.0, 0.70710678), flip_r: bool = False):
        p, q = (C.c_float * 4)(*pos), (C.c_float * 4)(*quat)
        _check(lib().grt_group_set_camera(self._g, p, q, int(flip_r)))

    def set_frames_in_flight(self, n: int):
        _check(lib().grt_group_set_frames_in_flight(self._g, n))

    def render(self, width: int, height: int, fmt: int = GRT_FORMAT_RGBA8, host_out=None):
        """Queue one frame on all devices (asynchronous); `host_out` (pinned) receives the assembled frame."""
        _check(lib().grt_group_render(self._g, width, height, fmt, _ptr(host_out) if host_out is not None else None))

    def frame_ptr(self) -> int:
        return lib().grt_group_frame(self._g) or 0

    def synchronize(self):
        _check(lib().grt_group_synchronize(self._g))


def interleaved_stripes(height: int, rank: int, world: int, stripe_rows: int = 54) -> List[Tuple[int, int]]:
    """Row stripes owned by `rank` when a frame is tiled over `world` GPUs: even-height stripes dealt
    round-robin so the expensive rows around the hole are shared out (grt_interleaved_stripes)."""
    cap = height // 2 + 2
    y0, y1 = (C.c_int * cap)(), (C.c_int * cap)()
    n = _check(lib().grt_interleaved_stripes(height, rank, world, stripe_rows, y0, y1, cap))
    return [(y0[i], y1[i]) for i in range(n)]


def balanced_stripes(height: int, rank: int, world: int, costs, stripe_rows: int = 54) -> List[Tuple[int, int]]:
    """Row stripes owned by `rank` when the stripes are dealt by measured cost, longest first (grt_balanced_stripes);
    `costs` = Context.measure_stripe_costs(...) of the frame, the same numbers on every rank."""
    cap = height // 2 + 2
    y0, y1 = (C.c_int * cap)(), (C.c_int * cap)()
    cost = (C.c_double * len(costs))(*[float(x) for x in costs])
    n = _check(lib().grt_balanced_stripes(height, rank, world, stripe_rows, cost, len(costs), y0, y1, cap))
    return [(y0[i], y1[i]) for i in range(n)]


def contiguous_bounds(height: int, world: int, costs=None, stripe_rows: int = 54) -> List[int]:
    """world + 1 row boundaries of equal-cost contiguous blocks (grt_contiguous_bounds); `costs` from Context.measure_stripe_costs
    or None for equal heights.  Rank r renders rows [bounds[r], bounds[r + 1])."""
    out = (C.c_int * (world + 1))()
    if costs is None:
        _check(lib().grt_contiguous_bounds(height, world, stripe_rows, None, 0, out))
    else:
        cost = (C.c_double * len(costs))(*[float(x) for x in costs])
        _check(lib().grt_contiguous_bounds(height, world, stripe_rows, cost, len(costs), out))
    return list(out)


def refine_bounds(height: int, bounds: List[int], measured, costs=None, stripe_rows: int = 54) -> List[int]:
    """Boundaries moved by the measured cost of every rank's block (grt_refine_bounds); `costs`: the stripe-cost profile the
    bounds came from (its shape is kept inside a block)."""
    world = len(bounds) - 1
    b = (C.c_int * (world + 1))(*bounds)
    m = (C.c_double * world)(*[float(x) for x in measured])
    out = (C.c_int * (world + 1))()
    if costs is None:
        _check(lib().grt_refine_bounds(height, world, stripe_rows, None, 0, b, m, out))
    else:
        cost = (C.c_double * len(costs))(*[float(x) for x in costs])
        _check(lib().grt_refine_bounds(height, world, stripe_rows, cost, len(costs), b, m, out))
    return list(out)


def save_png(path: str, rgba: np.ndarray) -> None:
    """HxWx4 uint8 -> PNG file (grt_image_save_png: the reference's screenshot file format, main.cpp:2783-2812)."""
    a = np.ascontiguousarray(rgba, np.uint8)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("expected an HxWx4 uint8 image")
    _check(lib().grt_image_save_png(path.encode(), a.ctypes.data, a.shape[1], a.shape[0]))


def load_png(path: str) -> np.ndarray:
    """PNG file -> HxWx4 uint8 (grt_image_load_png: the library's own decoder; the reference uses sf::Image)."""
    w, h = C.c_int(), C.c_int()
    p = lib().grt_image_load_png(os.fsencode(path), C.byref(w), C.byref(h))
    if not p:
        raise GrtError(-1, _err())
    try:
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(h.value, w.value, 4)).copy()
    finally:
        lib().grt_image_free(p)


def synthetic_background(width: int = 4096, height: int = 2048, seed: int = 7) -> np.ndarray:
    """Deterministic RGBA8 sky texture (graded bands + a lat/long grid + stars) standing in for the
    reference's 4096x2048 `backgrounds/nasa.png`, which is not redistributable with this repo."""
    rng = np.random.default_rng(seed)
    v = np.linspace(0.0, 1.0, height, dtype=np.float32)[:, None]
    u = np.linspace(0.0, 1.0, width, dtype=np.float32)[None, :]
    img = np.zeros((height, width, 4), dtype=np.float32)
    img[..., 0] = 0.25 + 0.25 * np.sin(2 * np.pi * (3 * u + v))
    img[..., 1] = 0.25 + 0.25 * np.sin(2 * np.pi * (5 * v - u) + 1.0)
    img[..., 2] = 0.35 + 0.30 * np.cos(2 * np.pi * (2 * u + 2 * v))
    grid = ((np.abs(((u * 24) % 1.0) - 0.5) > 0.47) | (np.abs(((v * 12) % 1.0) - 0.5) > 0.47))
    img[grid, :3] = 0.9
    n_stars = width * height // 400
    ys = rng.integers(0, height, n_stars)
    xs = rng.integers(0, width, n_stars)
    img[ys, xs, :3] = rng.uniform(0.6, 1.0, (n_stars, 1)).astype(np.float32)
    img[..., 3] = 1.0
    return (np.clip(img, 0, 1) * 255.0 + 0.5).astype(np.uint8)
