"""TEST INFRASTRUCTURE ONLY - recipe for oracle/_ref/: the reference's own code, built where it lies.

`build_clref(metric_source, name)` compiles /root/reference/cl.cl - the reference's OpenCL C device
source, all kernels of the hot path - as C++ for the host, for one metric:

    sed <token rewrites> /root/reference/cl.cl ; cat ref_shim/cl_trailer.inc      (stdin, nothing is copied)
      | g++ -x c++ -include ref_shim/opencl_as_cxx.hpp -include ref_shim/cl_macro_abi.hpp -I /root/reference -
    + ref_driver.cpp  ->  oracle/_ref/libclref_<name>_<hash>.so

The only textual rewrite is OpenCL's vector-literal cast `(float4)(a, b, c, d)` -> `float4(a, b, c, d)`
(in C++ the former is a cast of a comma expression); everything else is handled by the shim header.
The `-D` macro set cl.cl expects (metric.hpp:725-959) comes from ref_shim/cl_macro_abi.hpp over the
functions our emitter generated for the metric, because the reference's symbolic library (deps/vec)
is an empty submodule here.  Outputs go only into oracle/_ref/ (git-ignored, travels to the GPU
box).  On machines without /root/reference (the GPU box) a prebuilt library is used if present.
"""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
REFERENCE_ROOT = "/root/reference"
CL_SOURCE = os.path.join(REFERENCE_ROOT, "cl.cl")
SED_REWRITE = r"s/\((float[234]|int[234])\)\(/\1(/g"
_SHIM = [os.path.join(_HERE, "ref_shim", f) for f in ("opencl_as_cxx.hpp", "cl_macro_abi.hpp", "cl_trailer.inc")]
_DRIVER = [os.path.join(_HERE, f) for f in ("ref_driver.cpp", "geodesic_oracle.hpp")]


def reference_available() -> bool:
    return os.path.exists(CL_SOURCE)


def _source_tag(metric_source: str) -> str:
    """Hash of everything of OURS that goes into the library (so the prebuilt one is found without the reference)."""
    h = hashlib.sha1(metric_source.encode())
    for s in _SHIM + _DRIVER:
        h.update(open(s, "rb").read())
    return h.hexdigest()[:16]


def _cpu_has_v3() -> bool:
    try:
        flags = set([ln for ln in open("/proc/cpuinfo") if ln.startswith("flags")][0].split())
    except Exception:
        return False
    return {"avx2", "fma", "bmi2", "f16c", "movbe"} <= flags


def build_clref(metric_source: str, name: str, native: bool = False) -> str | None:
    """Returns the path of libclref for this metric, building it if the reference is mounted; None if neither
    a prebuilt library nor the reference exists.  `native` = the CPU-baseline build: -O3 for x86-64-v3 (AVX2 + FMA;
    not -march=native, because the library is built in one container and timed on another host)."""
    os.makedirs(REF_DIR, exist_ok=True)
    if native and not _cpu_has_v3():
        native = False
    tag = _source_tag(metric_source) + ("_v3u" if native else "")
    so = os.path.join(REF_DIR, f"libclref_{name}_{tag}.so")
    if os.path.exists(so):
        return so
    if not reference_available():
        if native:   # a generic prebuilt library is still the reference's code
            return build_clref(metric_source, name, native=False)
        return None
    hdr = os.path.join(REF_DIR, f"metric_{name}_{tag}.h")   # OUR generated metric functions (not reference text)
    with open(hdr, "w") as f:
        f.write(metric_source)
    # the timed build mirrors the reference's own compile options, "-cl-unsafe-math-optimizations" (metric_manager.hpp:70);
    # the pinning build stays strict IEEE so that it can be compared bit for bit with the restatement
    opt = ["-O3", "-march=x86-64-v3", "-funsafe-math-optimizations", "-fno-math-errno"] if native else ["-O2"]
    common = ["-std=c++20", "-fopenmp", "-fPIC", "-fpermissive", "-w", "-fvisibility=hidden", f'-DGRT_METRIC_HEADER="{hdr}"', "-I", _HERE] + opt
    obj_cl, obj_drv = so + ".cl.o", so + ".drv.o"
    pipeline = (f"(sed -E '{SED_REWRITE}' '{CL_SOURCE}'; cat '{_SHIM[2]}') | g++ -x c++ " + " ".join(f"'{c}'" for c in common) +
                f" -I '{REFERENCE_ROOT}' -include '{_SHIM[0]}' -include '{_SHIM[1]}' -c - -o '{obj_cl}'")
    subprocess.run(["bash", "-o", "pipefail", "-c", pipeline], check=True)
    subprocess.run(["g++"] + common + ["-fvisibility=default", "-c", _DRIVER[0], "-o", obj_drv], check=True)
    subprocess.run(["g++", "-shared", "-fopenmp", obj_cl, obj_drv, "-o", so + ".tmp"], check=True)
    os.replace(so + ".tmp", so)
    for f in (obj_cl, obj_drv):
        os.remove(f)
    return so


def prebuilt(name: str) -> list[str]:
    return sorted(glob.glob(os.path.join(REF_DIR, f"libclref_{name}_*.so")))


# ---- the reference's JavaScript engine (vendored QuickJS 2020-04-12 + bignum extension) -----------------------------------
QUICKJS_DIR = os.path.join(REFERENCE_ROOT, "deps", "quickjs")
NUMBER_JS = os.path.join(REFERENCE_ROOT, "number.js")
_QJS_SOURCES = ["cutils.c", "libbf.c", "libregexp.c", "libunicode.c", "quickjs.c"]


def build_qjsref() -> str | None:
    """oracle/_ref/libqjsref.so: QuickJS compiled from /root/reference/deps/quickjs where it lies (same defines as the
    reference's xmake.lua:45-46: CONFIG_BIGNUM, CONFIG_VERSION) + oracle/ref_quickjs.cpp (our numeric CShim/CMath binding).
    None when neither the library nor the reference is present."""
    os.makedirs(REF_DIR, exist_ok=True)
    binding = os.path.join(_HERE, "ref_quickjs.cpp")
    tag = hashlib.sha1(open(binding, "rb").read()).hexdigest()[:16]
    so = os.path.join(REF_DIR, f"libqjsref_{tag}.so")
    if os.path.exists(so):
        return so
    if not os.path.isdir(QUICKJS_DIR):
        return None
    objs = []
    for src in _QJS_SOURCES:
        obj = os.path.join(REF_DIR, f"qjs_{src}.o")
        subprocess.run(["gcc", "-O2", "-fPIC", "-w", "-D_GNU_SOURCE", "-DCONFIG_BIGNUM", '-DCONFIG_VERSION="2020-04-12"', "-c", os.path.join(QUICKJS_DIR, src), "-o", obj],
                       check=True)
        objs.append(obj)
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", QUICKJS_DIR, binding] + objs + ["-lm", "-ldl", "-lpthread", "-o", so + ".tmp"], check=True)
    os.replace(so + ".tmp", so)
    for o in objs:
        os.remove(o)
    return so
