"""Timeline of CTA 0 of k_tc_rowgemm (TENSORF_TC_TRACE=1): per-chunk clock64 stamps per role."""
import os, sys, ctypes, numpy as np, torch
os.environ["TENSORF_TC_TRACE"] = "1"
sys.path.insert(0, "tensorf-jax_b200")
from tensorf_b200 import _lib, ops
M, K, N, ns = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (156672, 160, 128, 3)))
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
A = torch.from_numpy(rng.normal(size=(M, K)).astype(np.float32)).to(dev)
W = torch.from_numpy((rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)).to(dev)
bias = torch.zeros(N, device=dev)
out = torch.zeros((M, N), dtype=torch.float32, device=dev)
scratch = torch.empty(8 << 20, dtype=torch.uint8, device=dev)
lib = _lib.load()
for rep in range(3):
    _lib.check(lib.tensorf_tc_rowgemm_test(ops._stream(), A.data_ptr(), M, K, W.data_ptr(), N, bias.data_ptr(), 1, None, None, out.data_ptr(), scratch.data_ptr(), scratch.numel(), ns))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
_lib.check(lib.tensorf_tc_rowgemm_test(ops._stream(), A.data_ptr(), M, K, W.data_ptr(), N, bias.data_ptr(), 1, None, None, out.data_ptr(), scratch.data_ptr(), scratch.numel(), ns))
e1.record(); torch.cuda.synchronize()
print("kernel+pack ms", e0.elapsed_time(e1))
buf = (ctypes.c_longlong * 8192)()
_lib.check(lib.tensorf_tc_trace_read(buf, 8192))
t = np.array(buf[:], dtype=np.int64).reshape(8, 1024)
nch = (K + 31) // 32
ntile = (M + 127) // 128; mine = (ntile + 147) // 148
nseq = nch * mine
base = t[0, 0]
names = ["P.start", "P.afterEmptyWait", "P.stored", "P.arrived", "M.fullWaitDone", "M.committed", "E.tfullWaitDone", "E.tileDone"]
print("chunks per tile", nch, "tiles for CTA0", mine)
print("seq  " + "  ".join(f"{n:>16s}" for n in names[:6]))
for q in range(min(nseq, 24)):
    print(f"{q:3d}  " + "  ".join(f"{int(t[s, q] - base):16d}" for s in range(6)))
print("tile  E.start  E.bulkReadDone  E.tfullWaitDone  E.loopDone  E.bulkIssued  E.tileDone")
for i in range(mine):
    print(i, *(int(x - base) for x in (t[4, i], t[4, 512 + i], t[6, i], t[5, i], t[5, 512 + i], t[7, i])))
print("total cycles (last tile done)", int(t[7, mine - 1] - base))

