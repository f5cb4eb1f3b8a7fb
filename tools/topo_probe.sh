nvidia-smi topo -m
lscpu | head -40
cat /proc/self/status | grep -i allowed
ls /sys/devices/system/node/
for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemTotal $n/meminfo); done
for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do b=$(echo $d | tr A-Z a-z | sed 's/^0000//'); echo $d $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) $(cat /sys/bus/pci/devices/$b/local_cpulist 2>/dev/null); done
nproc; free -g
which numactl
cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null
