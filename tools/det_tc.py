"""Run the row GEMM test entry repeatedly and count bitwise differences between runs (race detector)."""
import os, sys, numpy as np, torch
sys.path.insert(0, "tensorf-jax_b200")
from tensorf_b200 import _lib, ops
dev = torch.device("cuda:0")
lib = _lib.load()
rng = np.random.default_rng(0)
def run(M, K, N, ns, reps=8):
    A = torch.from_numpy(rng.normal(size=(M, K)).astype(np.float32)).to(dev)
    W = torch.from_numpy((rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)).to(dev)
    bias = torch.zeros(N, device=dev)
    scratch = torch.empty(8 << 20, dtype=torch.uint8, device=dev)
    outs = []
    for r in range(reps):
        out = torch.zeros((M, N), dtype=torch.float32, device=dev)
        _lib.check(lib.tensorf_tc_rowgemm_test(ops._stream(), A.data_ptr(), M, K, W.data_ptr(), N, bias.data_ptr(), 1, None, None, out.data_ptr(), scratch.data_ptr(), scratch.numel(), ns))
        torch.cuda.synchronize()
        outs.append(out)
    ref = (A.double() @ W.double()).clamp_min(0)
    d = [int((o != outs[0]).sum()) for o in outs[1:]]
    err = [float((o.double() - ref).abs().max()) for o in outs]
    bad_rows = torch.nonzero((outs[1] != outs[0]).any(dim=1)).flatten()[:12].tolist() if d[0] else []
    return d, max(err), bad_rows
for env in ({}, {"TENSORF_TC_RAW_SLOTS": "3"}, {"TENSORF_TC_STREAM_W": "1"}, {"TENSORF_TC_NO_TMA": "1"}):
    for k in ("TENSORF_TC_RAW_SLOTS", "TENSORF_TC_STAGES", "TENSORF_TC_NO_TMA", "TENSORF_TC_STREAM_W", "TENSORF_TC_DBG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for (M, K, N, ns) in ((155648, 128, 128, 2), (155648, 160, 128, 3), (155648, 128, 128, 3), (155648, 144, 32, 3)):
        print(env, (M, K, N, ns), *run(M, K, N, ns), flush=True)
