import sys, copy, yaml, numpy as np, torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from test_gpu_env import make_env
cfg = yaml.safe_load(open('envs/T1.yaml'))
# ---- terrain
env = make_env(cfg, 64, terrain="trimesh")
hf = env.terrain.height_field_raw
g = np.random.default_rng(0)
K = 200000
xy = np.stack([g.uniform(-4.9, 84.7, K), g.uniform(-4.9, 14.7, K)], axis=1).astype(np.float32)
xy[:1000] = (np.round(xy[:1000] * 10) / 10).astype(np.float32)
xy[1000:2000, 0] = g.uniform(-5.09, -5.0, 1000).astype(np.float32)
out = env.terrain.terrain_heights(torch.from_numpy(xy).cuda()).cpu().numpy()
bp, hs, vs = env.terrain.border_pixels, env.terrain.horizontal_scale, env.terrain.vertical_scale
x = bp + xy[:, 0] / hs; y = bp + xy[:, 1] / hs
x1 = np.floor(x).astype(int); x2 = x1 + 1; y1 = np.floor(y).astype(int); y2 = y1 + 1
ref = (((x2 - x) * (y2 - y) * hf[x1, y1] + (x - x1) * (y2 - y) * hf[x2, y1] + (x2 - x) * (y - y1) * hf[x1, y2] + (x - x1) * (y - y1) * hf[x2, y2]) * vs).astype(np.float32)
bad = np.nonzero(out.view(np.uint32) != ref.view(np.uint32))[0]
print('terrain mismatches', len(bad), bad[:10])
for i in bad[:8]:
    print(i, xy[i], x[i], y[i], out[i], ref[i], hf[x1[i], y1[i]], hf[x2[i], y1[i]], hf[x1[i], y2[i]], hf[x2[i], y2[i]])
env.close()
# ---- standing
for label, over in (("DR on", {}), ("DR off", "off")):
    c = copy.deepcopy(cfg)
    c["control"]["stiffness"] = {"Hip": 200.0, "Knee": 200.0, "Ankle": 400.0}
    c["control"]["damping"] = {"Hip": 5.0, "Knee": 5.0, "Ankle": 5.0}
    if over == "off":
        for k in list(c["randomization"].keys()):
            if isinstance(c["randomization"][k], dict):
                c["randomization"][k] = None
        c["noise"] = {}
    env = make_env(c, 512)
    obs, ex = env.reset()
    act = torch.zeros(512, 12, device='cuda')
    tot = 0
    for s in range(100):
        obs, rew, done, ex = env.step(act)
        tot += int(done.sum())
        if s % 10 == 0 or done.any():
            z = env.root_states[:, 2]
            v2 = env.root_states[:, 7:13].square().sum(1)
            print(label, s, 'z %.3f/%.3f' % (z.min().item(), z.mean().item()), 'vsq max %.2f' % v2.max().item(), 'done', int(done.sum()), 'pg z', env.projected_gravity[:, 2].mean().item(), 'ff', env._fview('feet_force').sum(1).mean().item())
        if tot > 600: break
    env.close()
