#!/bin/bash
nvidia-smi topo -m 2>&1 | head -20
lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name"
echo "cpuset: $(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null) mems: $(cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null)"
grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
which numactl; ls /usr/lib/x86_64-linux-gnu | grep -i numa
python - <<'P'
import os, pynvml
pynvml.nvmlInit()
n = pynvml.nvmlDeviceGetCount()
print("nvml gpus", n, "cpu_count", os.cpu_count(), "affinity", sorted(os.sched_getaffinity(0)))
for i in range(n):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    try:
        aff = pynvml.nvmlDeviceGetCpuAffinity(h, 8)
        print(i, "cpu affinity words", [hex(x) for x in aff])
    except Exception as e:
        print(i, "affinity err", e)
    try:
        print(i, "numa", pynvml.nvmlDeviceGetNumaNodeId(h))
    except Exception as e:
        print(i, "numa err", e)
P
