"""CPU experiment: how accurate must the MLP forward operands be?  Emulates split-precision
tensor-core GEMMs (fp16/bf16 hi+lo(+lo2) operands, fp32-ish accumulate) inside the fp64 oracle's
forward and reports the error of every gradient leaf against the exact fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tensorf-jax_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
import tensorf_oracle as O
from tensorf_b200 import synthetic as S
import helpers as H

def split(x, dt, n):
    parts, r = [], x.to(torch.float32)
    for _ in range(n):
        p = r.to(dt).to(torch.float32)
        parts.append(p); r = r - p
    return parts

class QMM(torch.autograd.Function):
    mode = ("fp16", 2)
    @staticmethod
    def forward(ctx, a, w):
        ctx.save_for_backward(a, w)
        dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[QMM.mode[0]]
        n = QMM.mode[1]
        A, W = split(a, dt, n), split(w, dt, n)
        out = torch.zeros(a.shape[0], w.shape[1], dtype=torch.float64)
        for i in range(n):
            for j in range(n):
                if i + j < n:   # keep terms down to the n-th order
                    out += A[i].double() @ W[j].double()
        return out.to(torch.float32).to(a.dtype)
    @staticmethod
    def backward(ctx, g):
        a, w = ctx.saved_tensors
        return g @ w.t(), a.t() @ g

def qmlp(cfg, mlp, features, viewdirs, camera_indices, aux=None):
    f = QMM.apply(features, mlp["w0"])
    x = torch.cat([f, viewdirs, O.fourier_encode(f, cfg.feature_n_freqs), O.fourier_encode(viewdirs, cfg.viewdir_n_freqs)], dim=-1)
    x = torch.relu(QMM.apply(x, mlp["w1"]) + mlp["b1"])
    x = torch.relu(QMM.apply(x, mlp["w2"]) + mlp["b2"])
    if cfg.num_cameras is not None:
        h = cfg.units // 2
        cond = mlp["embed"][camera_indices.to(torch.int64)]
        x = torch.cat([x[..., :h], cond[..., :h] * x[..., h:] + cond[..., h:]], dim=-1)
    return torch.sigmoid(x @ mlp["w3"] + mlp["b3"])

def run(w, R):
    inp = S.make_inputs(w, R=R)
    cfg, mc = H.oracle_cfgs(w)
    oi = H.oracle_inputs(inp, torch.float64)
    args = (cfg, mc, oi["params"], w.contracted, oi["aabb"], oi["origins"], oi["directions"], oi["camera_indices"], oi["colors"], oi["jitter"], oi["gumbel"])
    loss, rend, g_ref = O.loss_and_grads(*args)
    with torch.no_grad():
        _, aux = O.render_rays(*args[:8], args[9], args[10], return_aux=True)
    idx = aux["indices"]
    orig = O.feature_mlp
    for mode in [("bf16", 2), ("fp16", 2), ("bf16", 3)]:
        QMM.mode = mode
        O.feature_mlp = qmlp
        try:
            l2, r2, g2 = O.loss_and_grads(*args, forced_indices=idx)
        finally:
            O.feature_mlp = orig
        worst = 0
        rows = []
        for k in g_ref:
            a, b = g2[k].numpy(), g_ref[k].numpy()
            inf = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
            l2n = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
            rows.append((k, inf, l2n)); worst = max(worst, inf, l2n)
        print(w.name, R, mode, "rgb err %.2e" % float((r2 - rend).abs().max()), "worst grad rel %.2e" % worst)
        for k, inf, l2n in rows:
            if max(inf, l2n) > 2e-5: print("    %-20s inf %.2e l2 %.2e" % (k, inf, l2n))

if __name__ == "__main__":
    torch.set_num_threads(8)
    run(S.lego_workload(R=48, G=64, N=128, K=19), 48)
    run(S.dozer_workload(R=24, G=48, ncam=16), 24)
