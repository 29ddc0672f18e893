#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
run() { # tag set extra...
  local tag=$1 S=$2; shift 2
  local lib=""; [ "$tag" != main ] && lib="QT_LIB_PATH=$PWD/build_ab/$tag/libqtesla_b200.so"
  env $lib python bench.py --no-extras --set $S --steps 200 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$tag $S', round(d['value']/1e6,2), d['parity_check']['ok'])"
}
for t in main shift3 shift6 warps20; do run $t III; done
for t in main warps20; do run $t I; run $t p-I; done
for t in main nuss13; do run $t p-III; done
echo "== nuss F64 warps 12 (main) vs 13"
python tools/nuss_one.py III 1 3; python tools/nuss_one.py I 1 3
QT_LIB_PATH=$PWD/build_ab/nuss13/libqtesla_b200.so python tools/nuss_one.py III 1 3; QT_LIB_PATH=$PWD/build_ab/nuss13/libqtesla_b200.so python tools/nuss_one.py I 1 3
