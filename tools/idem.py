"""Forward twice (with a backward in between) and report whether rgb is bit-identical; env toggles bisect paths."""
import os, sys
sys.path.insert(0, "tensorf-jax_b200"); sys.path.insert(0, "tests"); sys.path.insert(0, "oracle"); sys.path.insert(0, ".")
import numpy as np, torch
from tensorf_b200 import ops, synthetic as S
from helpers import device_inputs
dev = torch.device("cuda:0")
w = S.lego_workload()
inp = S.make_inputs(w, bias_std=0.02)
params, dins = device_inputs(w, inp, dev)
desc = ops.make_desc(R=w.R, N=w.N, K=w.K, G=w.G, cd=w.cd, ca=w.ca, feat_freqs=w.feat_freqs, view_freqs=w.view_freqs, loss_scale=1.0 / (3 * w.R))
for envs in ({}, {"TENSORF_TC_NO_ENC_FUSION": "1"}, {"TENSORF_TC_NO_TMA": "1"}, {"TENSORF_TC_NO_TMA": "1", "TENSORF_TC_NO_ENC_FUSION": "1"}):
    for k in ("TENSORF_TC_NO_ENC_FUSION", "TENSORF_TC_NO_TMA"):
        os.environ.pop(k, None)
    os.environ.update(envs)
    call = ops.RenderCall(desc, dev)
    rgb, loss = call.forward(params, dins)
    rgb = rgb.clone()
    views = {n: call.view(n).clone() for n in ("rgb_sel", "feat")}
    call.backward()
    diffs = []
    for rep in range(3):
        rgb2, _ = call.forward(params, dins)
        diffs.append(int((rgb2 != rgb).sum()))
    v2 = {n: int((call.view(n) != views[n]).sum()) for n in views}
    print(envs, "rgb mismatches per rerun:", diffs, "views:", v2, flush=True)
