"""dev helper: clock64 timeline of one tile of update_wide_kernel (library built with GS_NVCC_EXTRA=-DGS_WIDE_TRACE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GS_DEV_TRACK"] = os.environ.get("GS_DEV_TRACK", "0")
os.environ["GS_DEV_HIDDEN"] = "256"
sys.argv = [sys.argv[0], "--child"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dev_update_time.py")).read())
import ctypes as C
buf = (C.c_longlong * 48)()
rc = C.CDLL(N.LIB_PATH).gs_debug_wide_trace(buf)
v = list(buf)
base = v[0]
n0 = ["A:wait z1", "A:start", "A:done", "B:start(z2 done)", "B:done", "C:out ready", "C:done", "D:start(dh2)", "D:done", "E:start(dh1)", "E:done"]
n1 = ["fwd:wait h1", "fwd:go", "-", "fwd:issued", "heads:go", "heads:issued", "dh2/Wc:go", "issued", "Wd/dgrad:go", "-", "dgrad issued", "Wb:go", "Wb issued", "L1 next issued"]
print("compute warp 0:  " + "  ".join(f"{n} {v[k]-base}" for k, n in enumerate(n0)))
print("MMA thread:      " + "  ".join(f"{n} {v[24+k]-base}" for k, n in enumerate(n1)))
