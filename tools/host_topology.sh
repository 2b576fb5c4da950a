#!/bin/bash
# Prints what the host exposes about CPU/NUMA/PCIe placement of the GPUs (diagnostic for the end-to-end mode).
nvidia-smi topo -m 2>&1 | head -40
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core"
ls /sys/devices/system/node/ 2>/dev/null | tr '\n' ' '; echo
for d in /sys/bus/pci/devices/*; do
  if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-6)" = "0x0302" ]; then
    echo "$(basename $d) numa_node=$(cat $d/numa_node) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null) max=$(cat $d/max_link_speed 2>/dev/null)"
  fi
done
nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current --format=csv
