import sys, os, numpy as np, torch, subprocess
sys.path.insert(0, "tensorf-jax_b200")
if len(sys.argv) > 1:
    M, K, N, ns = map(int, sys.argv[1:5])
    from tensorf_b200 import _lib, ops
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    A = torch.from_numpy(rng.normal(size=(M, K)).astype(np.float32)).to(dev)
    W = torch.from_numpy((rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)).to(dev)
    out = torch.zeros((M, N), dtype=torch.float32, device=dev)
    scratch = torch.empty(4 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.load()
    _lib.check(lib.tensorf_tc_rowgemm_test(ops._stream(), A.data_ptr(), M, K, W.data_ptr(), N, None, 0, None, None, out.data_ptr(), scratch.data_ptr(), scratch.numel(), ns))
    torch.cuda.synchronize()
    ref = A.double().cpu() @ W.double().cpu()
    print("OK rel", float((out.cpu().double() - ref).abs().max() / ref.abs().max()))
else:
    for env, cfgs in (({}, [(777, 390, 128, 3), (777, 384, 128, 3), (777, 260, 128, 3), (777, 200, 128, 3), (777, 390, 128, 2), (128, 390, 128, 3)]),
                      ({"TENSORF_TC_NOBULK": "1"}, [(777, 390, 128, 3)]), ({"TENSORF_TC_STAGES": "2"}, [(777, 390, 128, 3), (777, 150, 128, 3)])):
        for c in cfgs:
            r = subprocess.run([sys.executable, __file__] + [str(x) for x in c], env={**os.environ, **env}, capture_output=True, text=True, timeout=120)
            print(env, c, (r.stdout.strip().splitlines() or ["-"])[-1], "| rc", r.returncode, "|", (r.stderr.strip().splitlines() or [""])[-1][:100])
