"""Host-core discovery for the CPU oracle / reference timing (TEST INFRASTRUCTURE).

``os.cpu_count()`` reports the host's cores even when the container is limited by cpuset / cgroup quota;
running torch with that many threads oversubscribes the few cores actually granted (observed: 60 s instead of
1 s per ViT-B explanation on the GPU box).  ``usable_cpus()`` = min(affinity mask, cgroup quota)."""
import math
import os


def usable_cpus():
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    for path in ("/sys/fs/cgroup/cpu.max",):
        try:
            quota, period = open(path).read().split()[:2]
            if quota != "max":
                n = min(n, max(1, math.ceil(int(quota) / int(period))))
        except (OSError, ValueError):
            pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0 and p > 0:
            n = min(n, max(1, math.ceil(q / p)))
    except (OSError, ValueError):
        pass
    return max(1, n)


def set_torch_threads(cap=64):
    import torch
    n = min(usable_cpus(), cap)
    torch.set_num_threads(n)
    return n
