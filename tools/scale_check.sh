#!/bin/bash
# 1/2/4/8-GPU weak-scaling bench + strong-scaling batch sweep on one box (run with gpurun --gpus 8)
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then python bench.py --gpus 1 --steps 200 --warmup 10 --no-extras > $OUT/scale_n$N.json 2>/dev/null
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 200 --warmup 10 --no-extras > $OUT/scale_n$N.json 2>/dev/null; fi
  python - <<PY
import json
d=json.loads([l for l in open("$OUT/scale_n$N.json") if l.startswith("{")][-1]); print("N=$N", round(d["value"]/1e6,1), "M polymul/s", round(d["ms_per_step"],4), "ms/step", "e2e", round(d["e2e"]["value"]/1e6,2), "M/s", d["clocks"])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 tools/sweep.py > $OUT/sweep_8gpu.jsonl 2>/dev/null
python - <<PY
import json
for l in open("$OUT/sweep_8gpu.jsonl"):
    if not l.startswith("{"): continue
    d=json.loads(l); print(d["total_batch"], d["n_gpus"], round(d["us_per_step"],1), "us", round(d["polymuls_per_s"]/1e6,1), "M/s")
PY
