#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
val() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', round(d['value']/1e6,2), 'M polymul/s', d['parity_check']['ok'])"; }
for S in p-III; do python bench.py --no-extras --set $S --steps 100 > $OUT/ab_main_$S.json 2>/dev/null; val $OUT/ab_main_$S.json "main $S"; done
for S in p-III; do QT_LIB_PATH=$PWD/build_ab/fusedadd/libqtesla_b200.so python bench.py --no-extras --set $S --steps 100 > $OUT/ab_fa_$S.json 2>/dev/null; val $OUT/ab_fa_$S.json "fusedadd $S"; done
for S in p-III; do QT_LIB_PATH=$PWD/build_ab/fusedadd/libqtesla_b200.so python bench.py --no-extras --set $S --steps 100 --variant 3 > $OUT/ab_fa3_$S.json 2>/dev/null; val $OUT/ab_fa3_$S.json "fusedadd variant3 $S"; done
echo "== nuss: L2 prefetch (main) vs L1 prefetch"
python tools/nuss_one.py III 1 3; python tools/nuss_one.py III 0 0; python tools/nuss_one.py I 1 3
QT_LIB_PATH=$PWD/build_ab/pfl1/libqtesla_b200.so python tools/nuss_one.py III 1 3; QT_LIB_PATH=$PWD/build_ab/pfl1/libqtesla_b200.so python tools/nuss_one.py III 0 0; QT_LIB_PATH=$PWD/build_ab/pfl1/libqtesla_b200.so python tools/nuss_one.py I 1 3
QT_LIB_PATH=$PWD/build_ab/fusedadd/libqtesla_b200.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or fuzz or full_size or worst_case" 2>&1 | tail -2
