"""CPU ORACLE - TEST INFRASTRUCTURE ONLY (see geodesic_oracle.hpp).

Builds and binds (i) `Oracle`: the C++ restatement of the reference's OpenCL kernels for one metric, and
(ii) `ClReference`: the reference's OWN kernels, /root/reference/cl.cl compiled as C++ for the host into
oracle/_ref/ (ref_build.py).  (i) is pinned bit for bit against (ii) (tests/test_reference_kernels.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(_HERE, "_build")


class OracleFeatures(C.Structure):
    _fields_ = [
        ("universe_size", C.c_float), ("max_acceleration_change", C.c_float), ("max_precision_radius", C.c_float),
        ("min_step", C.c_float), ("adaptive_sampling_threshold", C.c_float), ("field_of_view", C.c_float),
        ("redshift", C.c_int), ("reparameterisation", C.c_int), ("use_old_redshift", C.c_int), ("adaptive_sampling", C.c_int),
    ]


class OracleFrameArgs(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("camera_pos", C.c_float * 4), ("camera_quat", C.c_float * 4),
        ("basis_speed", C.c_float * 3), ("flip", C.c_int), ("use_given_camera", C.c_int), ("camera_generic", C.c_float * 4),
        ("tetrad", C.c_float * 16), ("features", OracleFeatures), ("params", C.c_float * 256), ("max_probes", C.c_int),
        ("use_prepass", C.c_int), ("bg1", C.c_void_p), ("bg1_w", C.c_int), ("bg1_h", C.c_int), ("bg2", C.c_void_p),
        ("bg2_w", C.c_int), ("bg2_h", C.c_int),
    ]


class OracleFrameOut(C.Structure):
    _fields_ = [
        ("image", C.c_void_p), ("rays", C.c_void_p), ("rdata", C.c_void_p), ("termination", C.c_void_p),
        ("trial_steps", C.c_longlong), ("accepted_steps", C.c_longlong), ("rays_traced", C.c_longlong), ("extra_rays", C.c_longlong),
        ("t_init_s", C.c_double), ("t_integrate_s", C.c_double), ("t_render_data_s", C.c_double), ("t_adaptive_s", C.c_double),
        ("t_render_s", C.c_double), ("t_total_s", C.c_double), ("camera_generic", C.c_float * 4), ("tetrad", C.c_float * 16),
    ]


ORACLE_LIGHTRAY_DTYPE = np.dtype([
    ("position", "<f4", 4), ("velocity", "<f4", 4), ("initial_quat", "<f4", 4), ("acceleration", "<f4", 4),
    ("ku_uobsu", "<f4"), ("running_dlambda_dnew", "<f4"), ("terminated", "<i4"), ("sx", "<i4"), ("sy", "<i4"),
])
ORACLE_RENDER_DATA_DTYPE = np.dtype([
    ("tex_coord", "<f4", 2), ("z_shift", "<f4"), ("sx", "<i4"), ("sy", "<i4"), ("terminated", "<i4"), ("side", "<i4"),
])


def build(metric_source: str, name: str, native: bool = False) -> str:
    """Compile the oracle against one generated metric header. Returns the .so path (cached by hash)."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "geodesic_oracle.hpp")]
    h = hashlib.sha1(metric_source.encode())
    for s in srcs:
        h.update(open(s, "rb").read())
    if native:
        # -march=native objects must not travel between different host CPUs
        try:
            flags = [ln for ln in open("/proc/cpuinfo") if ln.startswith("flags")][0]
        except Exception:
            flags = "unknown"
        h.update(b"native" + flags.encode())
    else:
        h.update(b"generic")
    tag = h.hexdigest()[:16]
    so = os.path.join(BUILD_DIR, f"liboracle_{name}_{tag}.so")
    if os.path.exists(so):
        return so
    hdr = os.path.join(BUILD_DIR, f"metric_{name}_{tag}.h")
    with open(hdr, "w") as f:
        f.write(metric_source)
    cmd = ["g++", "-std=c++17", "-O2", "-fopenmp", "-shared", "-fPIC", f'-DGRT_METRIC_HEADER="{hdr}"', srcs[0], "-o", so + ".tmp"]
    if native:
        cmd[3:3] = ["-O3", "-march=native"]
    subprocess.run(cmd, check=True)
    os.replace(so + ".tmp", so)
    return so


class Oracle:
    _prefix = "oracle_"
    _rdata_dtype = ORACLE_RENDER_DATA_DTYPE
    _lightray_dtype = ORACLE_LIGHTRAY_DTYPE

    def __init__(self, metric_source: str, name: str, params: Sequence[float] = (), native: bool = False, threads: int = 0):
        """`threads` > 0 sets the OpenMP thread count explicitly (a launcher such as torchrun exports OMP_NUM_THREADS=1)."""
        self.path = self._build(metric_source, name, native)
        L = C.CDLL(self.path)
        self.lib = L
        for fn in ("info", "set_threads", "frame", "frame_window", "integrate_sample", "metric_eval", "tetrad_at", "geodesic_snapshot", "geodesic_interpolate"):   # one binding for both libraries
            setattr(L, "oracle_" + fn, getattr(L, self._prefix + fn))
        if threads:
            L.oracle_set_threads(int(threads))
        info = (C.c_int * 16)()
        L.oracle_info(info)
        self._info_raw = list(info)
        self.info = dict(zip(["big", "constant_theta", "singular", "adaptive_precision", "singularity_detection", "use_prepass", "num_params", "static_params",
                              "sizeof_lightray", "sizeof_render_data", "threads"], list(info)))
        assert self.info["sizeof_lightray"] == self._lightray_dtype.itemsize, self.info
        assert self.info["sizeof_render_data"] == self._rdata_dtype.itemsize, self.info
        self.params = np.zeros(256, np.float32)
        self.params[:len(params)] = params
        L.oracle_frame.argtypes = [C.POINTER(OracleFrameArgs), C.POINTER(OracleFrameOut)]
        L.oracle_frame_window.argtypes = [C.POINTER(OracleFrameArgs), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(OracleFrameOut)]
        L.oracle_integrate_sample.argtypes = [C.POINTER(OracleFrameArgs), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                              C.POINTER(C.c_double), C.c_void_p, C.c_void_p]

    @staticmethod
    def _build(metric_source, name, native):
        return build(metric_source, name, native)

    # -- helpers ---------------------------------------------------------------------------------
    @staticmethod
    def features(**kw) -> OracleFeatures:
        f = OracleFeatures(20.0, 0.01, 10.0, 1e-6, 32.0, 90.0, 0, 0, 0, 1)
        for k, v in kw.items():
            if not hasattr(f, k):
                raise AttributeError(k)
            setattr(f, k, v)
        return f

    def _args(self, width, height, features, camera_pos, camera_quat, basis_speed=(0, 0, 0), flip=False, max_probes=8, use_prepass=-1, bg1=None, bg2=None,
              given_camera=None) -> OracleFrameArgs:
        a = OracleFrameArgs()
        a.width, a.height = width, height
        a.camera_pos[:] = list(camera_pos)
        a.camera_quat[:] = list(camera_quat)
        a.basis_speed[:] = list(basis_speed)
        a.flip = int(flip)
        a.features = features
        a.params[:] = list(self.params)
        a.max_probes = max_probes
        a.use_prepass = use_prepass
        if given_camera is not None:
            a.use_given_camera = 1
            a.camera_generic[:] = list(given_camera[0])
            a.tetrad[:] = list(np.asarray(given_camera[1], np.float32).reshape(-1))
        self._keep = []
        if bg1 is not None:
            bg1 = np.ascontiguousarray(bg1, np.uint8)
            self._keep.append(bg1)
            a.bg1, a.bg1_w, a.bg1_h = bg1.ctypes.data, bg1.shape[1], bg1.shape[0]
        if bg2 is not None:
            bg2 = np.ascontiguousarray(bg2, np.uint8)
            self._keep.append(bg2)
            a.bg2, a.bg2_w, a.bg2_h = bg2.ctypes.data, bg2.shape[1], bg2.shape[0]
        return a

    def frame(self, width, height, features, camera_pos=(0, 0, -4, 0), camera_quat=(-0.70710678, 0, 0, 0.70710678), background=None, background2=None,
              want_rays=False, **kw):
        a = self._args(width, height, features, camera_pos, camera_quat, bg1=background, bg2=background2, **kw)
        o = OracleFrameOut()
        image = np.zeros((height, width, 4), np.float32) if background is not None else None
        rdata = np.zeros(width * height, self._rdata_dtype)
        rays = np.zeros(width * height, self._lightray_dtype) if want_rays else None
        pw, ph = width // 16, height // 16
        term = np.zeros(max(1, pw * ph), np.int32)
        o.image = image.ctypes.data if image is not None else None
        o.rdata = rdata.ctypes.data
        o.rays = rays.ctypes.data if rays is not None else None
        o.termination = term.ctypes.data
        rc = self.lib.oracle_frame(C.byref(a), C.byref(o))
        assert rc == 0
        return {
            "image": image, "rdata": rdata.reshape(height, width), "rays": rays, "termination": term[:pw * ph].reshape(ph, pw) if pw * ph else None,
            "trial_steps": o.trial_steps, "accepted_steps": o.accepted_steps, "rays_traced": o.rays_traced, "extra_rays": o.extra_rays,
            "t_integrate_s": o.t_integrate_s, "t_total_s": o.t_total_s, "t_render_s": o.t_render_s,
            "camera_generic": np.array(list(o.camera_generic), np.float32), "tetrad": np.array(list(o.tetrad), np.float32).reshape(4, 4),
        }

    def frame_window(self, width, height, window, features, camera_pos=(0, 0, -4, 0), camera_quat=(-0.70710678, 0, 0, 0.70710678), background=None, background2=None, **kw):
        """The window (x0, y0, x1, y1) of the width x height frame: same values as the full frame inside the window, but only the rays the
        window depends on are integrated (the 1/16 prepass runs in full).  Returns window-sized `image` and `rdata`."""
        x0, y0, x1, y1 = window
        a = self._args(width, height, features, camera_pos, camera_quat, bg1=background, bg2=background2, **kw)
        o = OracleFrameOut()
        ww, wh = x1 - x0, y1 - y0
        image = np.zeros((wh, ww, 4), np.float32) if background is not None else None
        rdata = np.zeros(ww * wh, self._rdata_dtype)
        o.image = image.ctypes.data if image is not None else None
        o.rdata = rdata.ctypes.data
        rc = self.lib.oracle_frame_window(C.byref(a), x0, y0, x1, y1, C.byref(o))
        assert rc == 0, rc
        return {"image": image, "rdata": rdata.reshape(wh, ww), "trial_steps": o.trial_steps, "rays_traced": o.rays_traced, "extra_rays": o.extra_rays,
                "t_integrate_s": o.t_integrate_s, "t_total_s": o.t_total_s,
                "camera_generic": np.array(list(o.camera_generic), np.float32), "tetrad": np.array(list(o.tetrad), np.float32).reshape(4, 4)}

    def integrate_sample(self, width, height, features, x0, y0, cw, ch, stride=1, camera_pos=(0, 0, -4, 0), camera_quat=(-0.70710678, 0, 0, 0.70710678)):
        a = self._args(width, height, features, camera_pos, camera_quat)
        trial, acc, sec = C.c_longlong(), C.c_longlong(), C.c_double()
        term = np.zeros(cw * ch, np.int32)
        steps = np.zeros(cw * ch, np.int32)
        self.lib.oracle_integrate_sample(C.byref(a), x0, y0, cw, ch, stride, C.byref(trial), C.byref(acc), C.byref(sec), term.ctypes.data, steps.ctypes.data)
        return {"trial_steps": trial.value, "accepted_steps": acc.value, "seconds": sec.value, "terminated": term.reshape(ch, cw), "steps": steps.reshape(ch, cw)}

    def metric_eval(self, pos):
        g = np.zeros(16, np.float32)
        dg = np.zeros(64, np.float32)
        p = np.ascontiguousarray(pos, np.float32)
        self.lib.oracle_metric_eval(p.ctypes.data_as(C.c_void_p), self.params.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), dg.ctypes.data_as(C.c_void_p))
        return g, dg

    def accel(self, pos, vel):
        p, v = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(vel, np.float32)
        out = [np.zeros(4, np.float32) for _ in range(3)]
        self.lib.oracle_accel(p.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), self.params.ctypes.data_as(C.c_void_p),
                              *[o.ctypes.data_as(C.c_void_p) for o in out])
        return {"ref": out[0], "numeric": out[1], "lowered": out[2]}

    def coords(self, pos, vel):
        p, v = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(vel, np.float32)
        outs = [np.zeros(4, np.float32) for _ in range(4)]
        dist = C.c_float()
        period = np.zeros(4, np.float32)
        self.lib.oracle_coords(p.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), self.params.ctypes.data_as(C.c_void_p),
                               *[o.ctypes.data_as(C.c_void_p) for o in outs], C.byref(dist), period.ctypes.data_as(C.c_void_p))
        return {"to_polar": outs[0], "to_polar_vel": outs[1], "from_polar": outs[2], "from_polar_vel": outs[3], "distance": dist.value, "period": period}

    def camera(self, cam_pos, flip=False, basis_speed=(0, 0, 0)):
        p = np.ascontiguousarray(cam_pos, np.float32)
        s = np.ascontiguousarray(basis_speed, np.float32)
        gen = np.zeros(4, np.float32)
        tet = np.zeros(16, np.float32)
        self.lib.oracle_camera(p.ctypes.data_as(C.c_void_p), int(flip), s.ctypes.data_as(C.c_void_p), self.params.ctypes.data_as(C.c_void_p),
                               gen.ctypes.data_as(C.c_void_p), tet.ctypes.data_as(C.c_void_p))
        return gen, tet.reshape(4, 4)

    def tetrad_at(self, generic_pos, basis_speed=(0, 0, 0), orient=False):
        p = np.ascontiguousarray(generic_pos, np.float32)
        s = np.ascontiguousarray(basis_speed, np.float32)
        tet = np.zeros(16, np.float32)
        self.lib.oracle_tetrad_at(p.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), int(orient), self.params.ctypes.data_as(C.c_void_p),
                                  tet.ctypes.data_as(C.c_void_p))
        return tet.reshape(4, 4)

    def mip_level(self, rgba, level):
        a = np.ascontiguousarray(rgba, np.uint8)
        lw, lh, levels = C.c_int(), C.c_int(), C.c_int()
        self.lib.oracle_mip_level(a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0], level, None, C.byref(lw), C.byref(lh), C.byref(levels))
        out = np.zeros((lh.value, lw.value, 4), np.uint8)
        self.lib.oracle_mip_level(a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0], level, out.ctypes.data_as(C.c_void_p), C.byref(lw), C.byref(lh), C.byref(levels))
        return out, levels.value

    def trace_path(self, pos, vel, features, max_points=20000):
        p, v = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(vel, np.float32)
        P = np.zeros((max_points, 4), np.float32)
        V = np.zeros((max_points, 4), np.float32)
        term = C.c_int()
        self.lib.oracle_trace_path.restype = C.c_int
        n = self.lib.oracle_trace_path(p.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.byref(features), self.params.ctypes.data_as(C.c_void_p), max_points,
                                       P.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p), C.byref(term))
        return P[:n], V[:n], term.value

    def geodesic_snapshot(self, camera_generic, tetrad, basis_speed, features, max_len=64000):
        cg = np.ascontiguousarray(camera_generic, np.float32)
        te = np.ascontiguousarray(tetrad, np.float32).reshape(-1)
        sp = np.ascontiguousarray(basis_speed, np.float32)
        pos = np.zeros((max_len, 4), np.float32)
        vel = np.zeros((max_len, 4), np.float32)
        ds = np.zeros(max_len, np.float32)
        tr = np.zeros((4, max_len, 4), np.float32)
        boosted = np.zeros(16, np.float32)
        self.lib.oracle_geodesic_snapshot.restype = C.c_int
        n = self.lib.oracle_geodesic_snapshot(cg.ctypes.data_as(C.c_void_p), te.ctypes.data_as(C.c_void_p), sp.ctypes.data_as(C.c_void_p), C.byref(features),
                                              self.params.ctypes.data_as(C.c_void_p), max_len, pos.ctypes.data_as(C.c_void_p), vel.ctypes.data_as(C.c_void_p),
                                              ds.ctypes.data_as(C.c_void_p), tr.ctypes.data_as(C.c_void_p), boosted.ctypes.data_as(C.c_void_p))
        return pos[:n], vel[:n], ds[:n], tr[:, :n], boosted.reshape(4, 4)

    def geodesic_interpolate(self, pos, vel, ds, tr, target_time, parallel_transport, basis_speed=(0, 0, 0)):
        n = len(pos)
        max_len = n
        pos = np.ascontiguousarray(pos, np.float32)
        vel = np.ascontiguousarray(vel, np.float32)
        ds = np.ascontiguousarray(ds, np.float32)
        tr = np.ascontiguousarray(tr, np.float32)
        sp = np.ascontiguousarray(basis_speed, np.float32)
        cam = np.zeros(4, np.float32)
        tet = np.zeros(16, np.float32)
        v = np.zeros(4, np.float32)
        self.lib.oracle_geodesic_interpolate(pos.ctypes.data_as(C.c_void_p), vel.ctypes.data_as(C.c_void_p), ds.ctypes.data_as(C.c_void_p), tr.ctypes.data_as(C.c_void_p),
                                             max_len, n, C.c_float(target_time), int(parallel_transport), sp.ctypes.data_as(C.c_void_p),
                                             self.params.ctypes.data_as(C.c_void_p), cam.ctypes.data_as(C.c_void_p), tet.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
        return cam, tet.reshape(4, 4), v


class ClReference(Oracle):
    """The reference's OWN kernels: /root/reference/cl.cl compiled as C++ for the host into oracle/_ref/
    (ref_build.py, ref_shim/).  Same calls as Oracle where the reference has the kernel: frame,
    integrate_sample, metric_eval, tetrad_at, plus step_verlet.  `accepted_steps` is -1 (the reference
    does not count them); `trial_steps` counts evaluations of its GEO_ACCEL0 macro."""
    _prefix = "clref_"
    # struct render_data of cl.cl:5066-5074 under OpenCL layout rules (float2 is 8-byte aligned): 32 bytes
    _rdata_dtype = np.dtype({"names": ["tex_coord", "z_shift", "sx", "sy", "terminated", "side"], "formats": [("<f4", 2), "<f4", "<i4", "<i4", "<i4", "<i4"],
                             "offsets": [0, 8, 12, 16, 20, 24], "itemsize": 32})
    # struct lightray of cl.cl:813-824: 84 bytes of fields, 16-byte aligned -> 96
    _lightray_dtype = np.dtype({"names": list(ORACLE_LIGHTRAY_DTYPE.names), "formats": [ORACLE_LIGHTRAY_DTYPE.fields[n][0] for n in ORACLE_LIGHTRAY_DTYPE.names],
                                "offsets": [ORACLE_LIGHTRAY_DTYPE.fields[n][1] for n in ORACLE_LIGHTRAY_DTYPE.names], "itemsize": 96})

    @staticmethod
    def _build(metric_source, name, native):
        from . import ref_build
        so = ref_build.build_clref(metric_source, name, native)
        if so is None:
            raise FileNotFoundError(f"no prebuilt oracle/_ref/libclref_{name}_*.so for this metric source and /root/reference is not mounted")
        return so

    def __init__(self, metric_source, name, params=(), native=False, threads=0):
        super().__init__(metric_source, name, params, native, threads)
        assert self._info_raw[11] == 1, "feature block layout mismatch"

    def step_verlet(self, pos, vel, acc, ds, features):
        """One trial step of the reference's step_verlet (cl.cl:3273-3346) -> (position, velocity, acceleration)."""
        p, v, a = (np.array(x, np.float32) for x in (pos, vel, acc))
        block = np.zeros(12, np.float32)      # struct dynamic_feature_config of the dynamic program: 7 floats, 5 ints
        block[:7] = [features.adaptive_sampling_threshold, features.field_of_view, features.max_acceleration_change, features.max_precision_radius,
                     features.min_step, 4.0, features.universe_size]
        block[7:].view(np.int32)[:] = [features.adaptive_sampling, features.redshift, features.reparameterisation, features.use_old_redshift, 0]
        self.lib.clref_step_verlet(p.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), C.c_float(ds),
                                   self.params.ctypes.data_as(C.c_void_p), block.ctypes.data_as(C.c_void_p))
        return p, v, a


def _stem(metric) -> str:
    return os.path.splitext(os.path.basename(metric.name.replace(" ", "_").replace("(", "").replace(")", "")))[0]


def for_metric(metric, is_static: bool = False, native: bool = False, threads: int = 0) -> Oracle:
    """Oracle for a geodesic_raytracing_b200.Metric (same generated functions as the CUDA program)."""
    return Oracle(metric.source(is_static), _stem(metric) + ("_static" if is_static else ""), metric.packed_params, native, threads)


def reference_for_metric(metric, native: bool = False, threads: int = 0) -> ClReference:
    """The reference's own cl.cl kernels (oracle/_ref) for a geodesic_raytracing_b200.Metric."""
    return ClReference(metric.source(False), _stem(metric), metric.packed_params, native, threads)


class QuickJSReference:
    """The reference's own JavaScript engine (vendored QuickJS, compiled from /root/reference/deps/quickjs into
    oracle/_ref/ by ref_build.build_qjsref) running number.js and a script, with CShim/CMath bound to numeric dual
    numbers (oracle/ref_quickjs.cpp).  `metric(js_path, pos)` -> g (16), dg (64, [k*16 + i*4 + j]) in float64."""

    def __init__(self):
        from . import ref_build
        so = ref_build.build_qjsref()
        if so is None or not os.path.exists(ref_build.NUMBER_JS):
            raise FileNotFoundError("QuickJS reference needs /root/reference (number.js, deps/quickjs)")
        self.number_js = ref_build.NUMBER_JS.encode()
        self.lib = C.CDLL(so)
        self.lib.qjsref_run.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                        C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.POINTER(C.c_double), C.c_char_p, C.c_int]

    def _run(self, path, kind, pos, params):
        names = list(params.keys())
        n = len(names)
        c_names = (C.c_char_p * max(1, n))(*[k.encode() for k in names])
        c_vals = (C.c_double * max(1, n))(*[float(params[k]) for k in names])
        total = {0: 16, 1: 4, 2: 1}[kind]
        out, dout = (C.c_double * total)(), (C.c_double * (total * 4))()
        out_len, n_names = C.c_int(), C.c_int()
        names_out = C.create_string_buffer(32 * 64)
        defaults = (C.c_double * 32)()
        err = C.create_string_buffer(1024)
        p = (C.c_double * 4)(*[float(x) for x in pos])
        rc = self.lib.qjsref_run(self.number_js, path.encode(), kind, p, n, c_names, c_vals, out, dout, C.byref(out_len), C.byref(n_names), names_out, defaults, err, 1024)
        found = {names_out.raw[i * 64:(i + 1) * 64].split(b"\0")[0].decode(): defaults[i] for i in range(n_names.value)}
        if rc != 0:
            raise RuntimeError(err.value.decode())
        return np.array(out[:]), np.array(dout[:]), out_len.value, found

    def run(self, path, kind, pos, params=None):
        """Two passes as the reference's sandbox implies: the first discovers `$cfg` names and `$default`s, the second
        evaluates with the defaults overridden by `params`."""
        _, _, _, found = self._run(path, kind, pos, {})
        values = dict(found)
        values.update(params or {})
        o, d, n, found2 = self._run(path, kind, pos, values)
        return o, d, n, found

    def metric(self, js_path, pos, params=None):
        g, dg, n, found = self.run(js_path, 0, pos, params)
        return g, dg, n, found
