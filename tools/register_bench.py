"""Timing of the registration step (match 1024 x 1024 x 32 descriptors + RANSAC over 10001 trials + refit) on one GPU,
next to the numpy oracle on the host."""
import importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
reg = importlib.import_module("3dfeatnet_b200.registration")
from oracle import registration as oreg

rng = np.random.default_rng(0)
n, d = 1024, 32
q = rng.normal(size=4); q /= np.linalg.norm(q)
R, t = oreg.quat2rot(q), rng.uniform(-5, 5, 3)
xyz2 = rng.uniform(-30, 30, (n, 3)).astype(np.float32); desc2 = rng.normal(size=(n, d)).astype(np.float32)
perm = rng.permutation(n)
xyz1 = (xyz2[perm] @ R.T + t).astype(np.float32); desc1 = desc2[perm] + rng.normal(0, 0.05, (n, d)).astype(np.float32)
bad = rng.random(n) < 0.7
desc1[bad] = rng.normal(size=(int(bad.sum()), d)).astype(np.float32)
dev = "cuda"
X1, D1, X2, D2 = (torch.as_tensor(a).to(dev) for a in (xyz1, desc1, xyz2, desc2))
triples = reg.draw_triples(n, 10001, seed=1)
tri = torch.as_tensor(triples).to(dev)
def gpu_pass():
    m = reg.match_descriptors(D1, D2)
    return reg.ransacfitRt(X1, X2[m.long()].contiguous(), 1.0, triples=tri.cpu().numpy())
gpu_pass(); torch.cuda.synchronize()
ts = []
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); s.record(); Rt, inl, trials = gpu_pass(); e.record(); torch.cuda.synchronize()
    ts.append((s.elapsed_time(e), (time.perf_counter() - t0) * 1e3))
# match kernel alone
m = reg.match_descriptors(D1, D2); torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); m = reg.match_descriptors(D1, D2); e.record(); torch.cuda.synchronize()
t0 = time.perf_counter()
rm, _ = oreg.match_descriptors(desc1, desc2)
ref = oreg.ransacfit_rt(xyz1, xyz2[rm], 1.0, triples)
cpu_ms = (time.perf_counter() - t0) * 1e3
print(json.dumps(dict(workload="register one pair: 1024 keypoints x 32-D, 70 % wrong descriptors, 10001 scored trials", trials_used=trials,
                      inliers=int(inl.numel()), gpu_ms_events=min(x[0] for x in ts), gpu_ms_wall=min(x[1] for x in ts), match_ms=s.elapsed_time(e),
                      cpu_oracle_ms=cpu_ms, cpu_trials=ref[2], same_inliers=bool(np.array_equal(inl.cpu().numpy(), ref[1])),
                      rot_err=float(np.abs(Rt.cpu().numpy()[:, :3] - R).max()))))
