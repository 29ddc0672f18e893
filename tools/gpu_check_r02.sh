#!/bin/bash
# One gpurun call (ONE GPU): tests, smoke, bench (both arms), Nussbaumer table, then the two ncu passes.
# Usage (from the repo root on the GPU box): bash tools/gpu_check_r02.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.csv 2>&1
nproc > $OUT/nproc_$TAG.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 2>/dev/null | grep -v '^count' > $OUT/bench_ref_$TAG.json; cut -c1-200 $OUT/bench_ref_$TAG.json
echo "== bench"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -3 $OUT/bench_$TAG.err; cut -c1-300 $OUT/bench_$TAG.json
echo "== nussbaumer table"; timeout 300 python tools/nuss_ab.py > $OUT/nuss_ab_$TAG.log 2>&1; grep -v whole $OUT/nuss_ab_$TAG.log
echo "== ncu launch list"
python bench.py --steps 5 --warmup 3 --no-extras > $OUT/ncu_plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/ncu_launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-extras > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
echo "== ncu full"
bash tools/ncu_capture.sh $TAG fused fusedpI fusedpIII fusedI nussF64
rm -f $OUT/ncu_fusedI_${TAG}_source.csv $OUT/ncu_nussF64_${TAG}_source.csv $OUT/ncu_fusedpIII_${TAG}_source.csv
ls -la $OUT | tail -30; du -sh $OUT
