import sys, torch
sys.path.insert(0, "tensorf-jax_b200")
from tensorf_b200 import _lib, ops
dev = torch.device("cuda:0")
out = torch.zeros(1, dtype=torch.int64, device=dev)
lib = _lib.load()
count = 256
for N in (32, 128, 256):
    for name, lt, lbo, sbo in (("none  KC32 (LBO128,SBO512)", 0, 128, 512), ("none  KC64 (LBO128,SBO1024)", 0, 128, 1024),
                               ("none  alt  (LBO2048,SBO128)", 0, 2048, 128),
                               ("SW32  (SBO256)", 6, 16, 256), ("SW64  (SBO512)", 4, 16, 512), ("SW128 (SBO1024)", 2, 16, 1024)):
        for rep in range(2):
            _lib.check(lib.tensorf_tc_umma_bench(ops._stream(), N, lt, lbo, sbo, count, out.data_ptr()))
            torch.cuda.synchronize()
        print(f"N={N:3d} {name:30s} cycles/MMA = {out.item() / count:7.1f}   (math floor {128 * N / 256:.0f})")
