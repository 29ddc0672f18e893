#!/bin/bash
# One gpurun call: tests, smoke, micro-benchmarks, bench (both arms), then the two ncu passes.
# Usage (from the repo root on the GPU box): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.csv 2>&1
nproc > $OUT/nproc_$TAG.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
echo "== ubench"; timeout 120 ./tools/ubench > $OUT/ubench_$TAG.json 2>&1; timeout 120 ./tools/synth > $OUT/synth_$TAG.json 2>&1; cat $OUT/synth_$TAG.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 2>/dev/null | grep -v '^count' > $OUT/bench_ref_$TAG.json; cut -c1-300 $OUT/bench_ref_$TAG.json
echo "== bench"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -3 $OUT/bench_$TAG.err; cut -c1-400 $OUT/bench_$TAG.json
echo "== sweep"; timeout 600 python tools/sweep.py > $OUT/sweep_1gpu_$TAG.jsonl 2>/dev/null; tail -3 $OUT/sweep_1gpu_$TAG.jsonl | cut -c1-300
echo "== nussbaumer table"; timeout 200 python tools/nuss_ab.py > $OUT/nuss_ab_$TAG.log 2>&1; cat $OUT/nuss_ab_$TAG.log
echo "== ncu launch list"
python bench.py --steps 5 --warmup 3 --no-extras > $OUT/ncu_plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-extras > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
echo "== ncu full (fused kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_polymul_tma -s 4 -c 1 -f -o $OUT/prof_fused_$TAG \
    python bench.py --steps 5 --warmup 3 --no-extras > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT | tail -20
