#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ (run from the repo root).

Source of truth: the fp64 instance of the CPU oracle (oracle/tensorf_oracle.py).  The reference
itself (brentyi/tensorf-jax) cannot be imported here — jax/flax are not installable in this
image — so these vectors pin the ORACLE (and through it the CUDA path) against regressions; they
are not outputs of the reference.  Inputs are regenerated from seeds by
tensorf_b200.synthetic.make_inputs, only outputs are stored (small files).
"""
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[2]
for p in (ROOT / "oracle", ROOT / "tests", ROOT / "tensorf-jax_b200"):
    sys.path.insert(0, str(p))

import tensorf_oracle as O  # noqa: E402
from helpers import oracle_cfgs, oracle_inputs  # noqa: E402
from tensorf_b200 import synthetic as S  # noqa: E402

CASES = {
    "lego_small": S.Workload("lego_small", 64, 9, 2, 3, 37, 5, 2, 2),
    "dozer_small": S.Workload("dozer_small", 32, 9, 4, 3, 40, 6, 6, 6, contracted=True, num_cameras=7),
}


def generate(name, w):
    inp = S.make_inputs(w, bias_std=0.05)
    cfg, mc = oracle_cfgs(w)
    oi = oracle_inputs(inp, torch.float64)
    loss, rgb, grads = O.loss_and_grads(cfg, mc, oi["params"], w.contracted, oi["aabb"], oi["origins"], oi["directions"],
                                        oi["camera_indices"], oi["colors"], oi["jitter"], oi["gumbel"])
    _, aux = O.render_rays(cfg, mc, oi["params"], w.contracted, oi["aabb"], oi["origins"], oi["directions"],
                           oi["camera_indices"], oi["jitter"], oi["gumbel"], return_aux=True)
    out = {"loss": loss.numpy(), "rgb": rgb.numpy(), "indices": np.sort(aux["indices"].numpy(), -1).astype(np.int32),
           "z": aux["z"].numpy(), "p_terminates": aux["p_terminates"].numpy()}
    for mode, key in ((O.DIST_MEDIAN, "depth_median"), (O.DIST_MEAN, "depth_mean")):
        c = O.RenderConfig(w.near, w.far, mode, w.N, w.K)
        out[key] = O.render_rays(c, mc, oi["params"], w.contracted, oi["aabb"], oi["origins"], oi["directions"],
                                 oi["camera_indices"], oi["jitter"], None).numpy()
    for k, g in grads.items():
        out["grad_" + k] = g.numpy()
    np.savez_compressed(pathlib.Path(__file__).parent / f"{name}.npz", **out)
    print(name, {k: v.shape for k, v in out.items() if k in ("rgb", "indices", "grad_density_matrix")}, float(loss))


if __name__ == "__main__":
    for n, w in CASES.items():
        generate(n, w)
