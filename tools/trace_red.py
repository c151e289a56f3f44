"""Timeline of CTA 0 of k_tc_redgemm (TENSORF_TC_TRACE=1)."""
import os, sys, ctypes, numpy as np, torch
os.environ["TENSORF_TC_TRACE"] = "1"
sys.path.insert(0, "tensorf-jax_b200")
from tensorf_b200 import _lib, ops
rows, Mg, Nx = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (156672, 128, 128)))
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
G = torch.from_numpy(rng.normal(size=(rows, Mg)).astype(np.float32)).to(dev)
X = torch.from_numpy(rng.normal(size=(rows, Nx)).astype(np.float32)).to(dev)
out = torch.zeros((Nx, Mg), dtype=torch.float32, device=dev)
lib = _lib.load()
for rep in range(3):
    _lib.check(lib.tensorf_tc_redgemm_test(ops._stream(), G.data_ptr(), Mg, X.data_ptr(), Nx, rows, out.data_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
_lib.check(lib.tensorf_tc_redgemm_test(ops._stream(), G.data_ptr(), Mg, X.data_ptr(), Nx, rows, out.data_ptr()))
e1.record(); torch.cuda.synchronize()
print("kernel ms", e0.elapsed_time(e1))
buf = (ctypes.c_longlong * 8192)()
_lib.check(lib.tensorf_tc_trace_read(buf, 8192))
t = np.array(buf[:], dtype=np.int64).reshape(8, 1024)
base = t[0, 0]
names = ["C.start", "C.rawReady", "C.stageFree", "C.arrived", "L.slotFree", "M.fullDone"]
print("chunk " + " ".join(f"{n:>14s}" for n in names))
for q in range(34):
    print(f"{q:3d}   " + " ".join(f"{int(t[s, q] - base):14d}" for s in range(6)))
print("E.tfull", int(t[6, 0] - base), "end", int(t[7, 0] - base))
