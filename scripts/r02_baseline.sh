#!/bin/bash
# round-2 baseline on the GPU box: topology for the NUMA work, a compute-sanitizer attempt, ncu captures of the sweep kernels
set -u
OUT=gpurun_out; mkdir -p $OUT
{ nvidia-smi topo -m; nvidia-smi -L; lscpu | head -30; numactl -H 2>&1 | head -20; ls /sys/devices/system/node/ 2>&1;
  for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done;
  ls /usr/lib/x86_64-linux-gnu | grep -i numa; python -c "import os; print('affinity', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8])"; nproc; free -g | head -2; } > $OUT/r02_topology.txt 2>&1
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_target.py > $OUT/r02_sanitizer.log 2>&1; echo "sanitizer rc=$?" >> $OUT/r02_sanitizer.log
python scripts/sweep_profile_target.py 22 > $OUT/r02a_sweeps_plain.log 2>&1 || { echo plain failed; tail -5 $OUT/r02a_sweeps_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'g1_smul|pairing_kernel|poly_mul_kernel|kzg_commit_kernel|poly_unary|ntt4_kernel' -c 40 -f -o $OUT/r02a_sweeps python scripts/sweep_profile_target.py 22 > $OUT/r02a_sweeps_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 $OUT/r02_sanitizer.log; cat $OUT/r02_topology.txt | head -60
