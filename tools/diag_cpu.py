import os, time, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(p, open(p).read().strip())
    except Exception as e: print(p, "n/a")
import torch
print("torch threads default", torch.get_num_threads())
from oracle import dinov2_restated as dr
t = time.time(); m = dr.build("dinov2_vitl14", seed=0, depth_override=4); print("build vitl d4 %.1fs" % (time.time() - t))
img = torch.randn(1, 3, 322, 322)
for nt in (8, 16, 32, 64, 128):
    torch.set_num_threads(nt)
    with torch.no_grad():
        m(img); t = time.time(); m(img); m(img); print(nt, "threads: fwd %.3fs" % ((time.time() - t) / 2))
