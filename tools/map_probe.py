import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from liodom_b200 import api, synth
ctx = api.Context(max_points=131072)
clouds, poses = [], []
T0 = synth.gt_pose(1000, 0, traj=1)
for k in range(200):
    f = k*10
    clouds.append(ctx.extract(synth.scan("hdl64_small", 1000, f, traj=1)))
    poses.append(np.linalg.inv(T0) @ synth.gt_pose(1000, f, traj=1))
ctx.close()
gm = api.Map(20.0, 25.0, 0.4, max_points=1<<22)
gm.update(clouds[0], poses[0])
tu = tl = 0.0
for c, T in zip(clouds, poses):
    t0=time.perf_counter(); gm.update(c, T); t1=time.perf_counter(); gm.get_local_map(T, 2, 1); t2=time.perf_counter()
    tu += t1-t0; tl += t2-t1
print("update ms/frame %.3f  get_local ms/frame %.3f  map points %d" % (tu/200*1e3, tl/200*1e3, gm.size()[0]))
