nvidia-smi topo -m 2>&1 | head -20
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"
echo affinity: $(taskset -p $$ 2>&1)
nproc
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo $d $(cat $d/numa_node); fi; done
cat /sys/devices/system/node/node*/cpulist 2>/dev/null
grep -i "mems_allowed\|cpus_allowed_list" /proc/self/status
free -g | head -2
