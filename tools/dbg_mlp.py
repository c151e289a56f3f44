import sys, numpy as np, torch
sys.path.insert(0, "tensorf-jax_b200")
from tensorf_b200 import ops, synthetic as S
dev = torch.device("cuda:0")
ca, F, V, ncam, rays, rpr = 48, 2, 2, None, 300, 33
M = rays * rpr
p_np = S.make_params(4, 1, ca, F, V, ncam, seed=11, bias_std=0.1)
rng = np.random.default_rng(7)
feat = torch.from_numpy(rng.normal(0, 0.3, (M, 3 * ca)).astype(np.float32)).to(dev)
vd = rng.normal(size=(rays, 3)).astype(np.float32); vd /= np.linalg.norm(vd, axis=-1, keepdims=True)
vd = torch.from_numpy(vd).to(dev)
d_rgb = torch.from_numpy(rng.normal(size=(M, 3)).astype(np.float32)).to(dev)
params = {k: torch.from_numpy(v).to(dev) for k, v in p_np.items()}
res = {}
for impl in (1, 2):
    desc = ops.make_desc(R=rays, N=rpr, K=rpr, G=4, cd=1, ca=ca, feat_freqs=F, view_freqs=V, num_cameras=ncam, mlp_impl=impl)
    call = ops.MlpCall(desc, M, dev)
    rgb = call.forward(params, feat, vd, None, rpr)
    d_feat, grads = call.backward(d_rgb)
    torch.cuda.synchronize()
    res[impl] = (rgb.cpu().numpy(), d_feat.cpu().numpy(), {k: v.cpu().numpy() for k, v in grads.items()})
a, b = res[1][1], res[2][1]
err = np.abs(a - b)
print("d_feat max", np.abs(a).max(), "max err", err.max(), "at", np.unravel_index(err.argmax(), err.shape))
bad = err > 1e-4 * np.abs(a).max()
print("bad count", bad.sum(), "of", bad.size)
rows = np.where(bad.any(axis=1))[0]; cols = np.where(bad.any(axis=0))[0]
print("bad rows", rows[:40], len(rows)); print("bad cols", cols[:40], len(cols))
print("row%128 of bad rows", sorted(set((rows % 128).tolist()))[:40])
for k in res[1][2]:
    e = np.abs(res[1][2][k] - res[2][2][k]).max() / (np.abs(res[1][2][k]).max() + 1e-30)
    print(k, "rel", e)
print("rgb rel", np.abs(res[1][0] - res[2][0]).max())
