"""Dev tool (GPU box): configs[2] family (2D-Laplacian pattern, 64-bit entries) at the given grid sides.

    python tools/lap_gpu.py 40 50 64
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

entry.build()
import slip_lu_b200  # noqa: E402

lib = slip_lu_b200.lib()
lib.dll.SLIP_B200_last_stats.argtypes = [C.POINTER(C.c_double), C.c_int]
lib.dll.SLIP_B200_last_pinv.argtypes = [C.POINTER(C.c_int32), C.c_int]
for m in [int(a) for a in sys.argv[1:]]:
    t = time.time()
    r = bench.laplacian_block(lib, m, 6458.4, "MEASURED_PEAKS.json")
    print(json.dumps(r), f"(tool wall {time.time() - t:.1f}s)", flush=True)
