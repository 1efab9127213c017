"""Compile a bundled metric's program here (no GPU) and print the static step analysis of every new cubin.
   python tools/compile_and_count.py [metric]; GRT_TORCH_NVRTC=1 loads PyTorch first (NVRTC 12.8 instead of the toolkit's 12.9)."""
import os, sys
if os.environ.get('GRT_TORCH_NVRTC'):
    import torch  # noqa: F401  (its bundled NVRTC 12.8 gets loaded before libgrt's)
import geodesic_raytracing_b200 as grt, time, glob
name = sys.argv[1] if len(sys.argv) > 1 else "kerr_schild"
c = grt.Content(grt.CONTENT_DIR)
m = c.load_by_name(name)
t0=time.time()
before=set(glob.glob("geodesic_raytracing_b200/cubin_cache/*.cubin"))
m.compile(is_static=False, fast_math=True)
after=set(glob.glob("geodesic_raytracing_b200/cubin_cache/*.cubin"))
for f in sorted(after-before, key=os.path.getmtime):
    hdr=open(f[:-6]+".cu").read(200).splitlines()[:2]
    print(f, hdr)
    os.system(f"strings {f}.log | grep -A3 'Compiling entry function .k_integrate' | grep -E 'Used|spill'; python tools/sass_step_static.py {f} | tail -3")
