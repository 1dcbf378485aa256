"""A worker that talks to the runtime ONLY through ctypes and the C ABI (no torch, no pybind): what a foreign binding would do.
Run under `kungfu-run -np N` (reference: kungfu/loader.py + the cgo exports of libkungfu-comm)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util

spec = importlib.util.spec_from_file_location("kf_loader", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kungfu_b200", "loader.py"))
loader = importlib.util.module_from_spec(spec)
spec.loader.exec_module(loader)  # loaded by path: importing the package would import torch
assert "torch" not in sys.modules

lib = loader.load_capi()
assert lib.kungfu_init() == 0, lib.kungfu_last_error()
n, r = lib.kungfu_size(), lib.kungfu_rank()
N = 100000
send = (ctypes.c_float * N)(*([float(r + 1)] * N))
recv = (ctypes.c_float * N)()
assert lib.kungfu_all_reduce(send, recv, N, loader.F32, loader.SUM, b"ctypes-sum", None, None) == 0
assert recv[0] == n * (n + 1) / 2 and recv[N - 1] == recv[0]
fired = []
cb = loader.CALLBACK(lambda arg: fired.append(1))
assert lib.kungfu_all_reduce(send, recv, N, loader.F32, loader.MAX, b"ctypes-max", ctypes.cast(cb, ctypes.c_void_p), None) == 0
assert lib.kungfu_barrier() == 0
import time

for _ in range(2000):
    if fired:
        break
    time.sleep(0.001)
assert fired and recv[0] == float(n)
v = ctypes.c_int64(r * 7)
assert lib.kungfu_all_reduce_int_max(ctypes.byref(v)) == 0 and v.value == (n - 1) * 7
blob = (ctypes.c_ubyte * 16)(*range(r, r + 16))
assert lib.kungfu_save(b"blob", blob, 16) == 0 and lib.kungfu_barrier() == 0
got, ok = (ctypes.c_ubyte * 16)(), ctypes.c_int(0)
assert lib.kungfu_request((r + 1) % n, b"", b"blob", got, 16, ctypes.byref(ok)) == 0 and ok.value == 1 and got[0] == (r + 1) % n
assert lib.kungfu_request((r + 1) % n, b"", b"nothing", got, 16, ctypes.byref(ok)) == 0 and ok.value == 0
assert loader._call_method(lib, "kungfu_barrier") and not loader._call_method(lib, "kungfu_no_such_symbol")
lib.kungfu_barrier()
print(json.dumps({"ok": True, "ctypes": True, "rank": r, "np": n, "uid": int(lib.kungfu_uid()) != 0}), flush=True)
lib.kungfu_finalize()
