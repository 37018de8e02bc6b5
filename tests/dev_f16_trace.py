"""dev helper: clock64 timeline of one tile of the fp16x3 update kernel (library built with GS_NVCC_EXTRA=-DGS_F16_TRACE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GS_DEV_TRACK"] = os.environ.get("GS_DEV_TRACK", "0")
sys.argv = [sys.argv[0], "--child"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dev_update_time.py")).read())
import ctypes as C
buf = (C.c_longlong * 64)()
rc = C.CDLL(N.LIB_PATH).gs_debug_f16_trace(buf)
v = list(buf)
names0 = ["A:wait z1", "A:start", "B:wait z2", "B:start", "C:enter", "C:start(loss)", "D:wait dh2", "D:start", "E:wait dh1", "E:start", "E:done", "C:ld done", "C:loss done", "C:g16 stored", "C:prefetch issued", "C:ready"]
names1 = ["L1 issue", "L1 issued", "fwd go", "fwd issued", "heads go", "heads issued", "dh2/Wc go", "issued", "dgrad/Wa go", "issued", "Wb go", "Wb issued"]
for s in range(2):
    base = v[s * 32]
    print(f"set {s}: compute warp 0 (cycles from its tile start)")
    print("   " + "  ".join(f"{n} {v[s*32+k]-base}" for k, n in enumerate(names0)))
    print(f"set {s}: MMA warp")
    print("   " + "  ".join(f"{n} {v[s*32+16+k]-base}" for k, n in enumerate(names1)))

tb = (C.c_longlong * 128)()
C.CDLL(N.LIB_PATH).gs_debug_f16_tiles(tb)
t = list(tb)
for s in range(2):
    t0 = t[s * 64 + 62]
    starts = [x - t0 for x in t[s * 64: s * 64 + 61] if x > 0]
    print(f"set {s}: kernel entry 0, first tile start {starts[0]}, tile durations {[b - a for a, b in zip(starts, starts[1:])]}, loop end {t[s*64+61]-t0}, exit {t[s*64+63]-t0}")
