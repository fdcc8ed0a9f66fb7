#!/bin/bash
# host / GPU topology of the box (for the e2e scaling analysis)
nvidia-smi topo -m 2>&1 | head -30
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302\|^0x0300" $d/class 2>/dev/null && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$(basename $d) numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist)"; fi; done
for n in /sys/devices/system/node/node*; do echo "$(basename $n) cpulist=$(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
grep -i "allowed_list" /proc/self/status
nproc
