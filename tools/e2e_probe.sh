#!/bin/bash
# Copy ceiling of the box vs the end-to-end hot path, for 1..N GPUs (run with gpurun --gpus N).
# Usage: bash tools/e2e_probe.sh <tag> [max_gpus] [full]   (full: also write-combined / NUMA-bound variants)
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l); MAXG=${2:-$NG}; FULL=${3:-}
{
  echo "== nvidia-smi topo -m"; nvidia-smi topo -m
  echo "== lscpu"; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core"
  echo "== numa"; ls /sys/devices/system/node/ | tr '\n' ' '; echo; cat /sys/devices/system/node/node*/cpulist 2>/dev/null
  echo "== mem"; free -g | head -2
  echo "== pcie link"; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
} > $OUT/topology_$TAG.txt 2>&1
timeout 300 ./tools/pcie_ceiling $MAXG $FULL > $OUT/pcie_ceiling_threads_$TAG.jsonl 2> $OUT/pcie_ceiling_threads_$TAG.err
: > $OUT/pcie_probe_$TAG.jsonl
for N in 1 2 4 8; do
  [ $N -le $MAXG ] || continue
  for FLAG in "" ${FULL:+--bind}; do
    if [ $N -eq 1 ]; then timeout 300 python tools/pcie_probe.py $FLAG --tag $TAG >> $OUT/pcie_probe_$TAG.jsonl 2>> $OUT/pcie_probe_$TAG.err
    else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+N)) \
           tools/pcie_probe.py $FLAG --tag $TAG >> $OUT/pcie_probe_$TAG.jsonl 2>> $OUT/pcie_probe_$TAG.err; fi
  done
done
grep -h '"summary"' $OUT/pcie_probe_$TAG.jsonl
cat $OUT/topology_$TAG.txt | head -40
grep -h '"copy"' $OUT/pcie_ceiling_threads_$TAG.jsonl | cut -c1-220
