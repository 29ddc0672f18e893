#!/bin/bash
# hi*q as shift-adds on every k-th butterfly (QT_SHIFT_MOD) re-measured on the final n=1024 kernel, run r02I
run() { local tag=$1 S=$2; shift 2
  local lib=""; [ "$tag" != main ] && lib="QT_LIB_PATH=$PWD/build_ab/$tag/libqtesla_b200.so"
  env $lib python bench.py --no-extras --set $S --steps 200 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$tag $S', round(d['value']/1e6,2), d['parity_check']['ok'])"
}
for t in main sm0 sm2 sm3 sm6 main sm0 sm3 sm6; do run $t III; done
