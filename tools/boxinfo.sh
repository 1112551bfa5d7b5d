#!/bin/bash
# host / topology facts of the GPU box (for the e2e host path): written to gpurun_out/boxinfo.txt
{
echo "== nproc / affinity"; nproc; python -c "import os;print(sorted(os.sched_getaffinity(0)))"
echo "== lscpu"; lscpu | head -40
echo "== numa nodes"; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemTotal $n/meminfo); done
echo "== numactl"; which numactl && numactl -H
echo "== topo"; nvidia-smi topo -m
echo "== gpus"; nvidia-smi --query-gpu=index,pci.bus_id,name,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do dd=$(echo $d | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo $d numa $(cat /sys/bus/pci/devices/$dd/numa_node 2>/dev/null); done
echo "== cgroup"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null
echo "== mem"; free -g | head -3
echo "== ulimit -l"; ulimit -l
} > gpurun_out/boxinfo.txt 2>&1
