#!/bin/bash
# run r02g: tests, then A/B of the p-I range plan, the bulk-store variant, the Nussbaumer L2 prefetch, the shuffle levels
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
val() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', round(d['value']/1e6,2), 'M polymul/s', round(d['ms_per_step'],4), 'ms', d['parity_check']['ok'], d['clocks']['sm_mhz'])"; }
for S in III I p-I p-III; do python bench.py --no-extras --set $S --steps 100 > $OUT/ab_main_$S.json 2>/dev/null; val $OUT/ab_main_$S.json "main $S"; done
for S in III I; do QT_LIB_PATH=$PWD/build_ab/tmastore/libqtesla_b200.so python bench.py --no-extras --set $S --steps 100 > $OUT/ab_tmastore_$S.json 2>/dev/null; val $OUT/ab_tmastore_$S.json "tmastore $S"; done
echo "== nuss prefetch (main) vs none"
python tools/nuss_one.py III 1 3; python tools/nuss_one.py III 0 0; python tools/nuss_one.py I 1 3; python tools/nuss_one.py p-I 1 0
QT_LIB_PATH=$PWD/build_ab/noprefetch/libqtesla_b200.so python tools/nuss_one.py III 1 3; QT_LIB_PATH=$PWD/build_ab/noprefetch/libqtesla_b200.so python tools/nuss_one.py III 0 0
QT_LIB_PATH=$PWD/build_ab/noprefetch/libqtesla_b200.so python tools/nuss_one.py I 1 3; QT_LIB_PATH=$PWD/build_ab/noprefetch/libqtesla_b200.so python tools/nuss_one.py p-I 1 0
echo "== shuffle levels vs transposition"; ./tools/shuffle_ab | tee $OUT/shuffle_ab_r02g.json
bash tools/ncu_capture.sh r02g fusedpI fused
