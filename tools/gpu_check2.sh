#!/bin/bash
# tests + ubench + bench (no ncu)
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu_$TAG.log
echo "== ubench"; timeout 120 ./tools/ubench > $OUT/ubench_$TAG.json 2>&1; cat $OUT/ubench_$TAG.json
echo "== bench"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -5 $OUT/bench_$TAG.err; cat $OUT/bench_$TAG.json
