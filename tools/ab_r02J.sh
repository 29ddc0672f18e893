#!/bin/bash
# QT_SHIFT_MOD 3 / 5 against 4 on every qTESLA-III kernel that shares the butterfly (fused, cached product, single transforms), run r02J
run() { local tag=$1 S=$2; shift 2
  local lib=""; [ "$tag" != main ] && lib="QT_LIB_PATH=$PWD/build_ab/$tag/libqtesla_b200.so"
  env $lib python bench.py --no-extras --set $S --steps 200 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$tag $S', round(d['value']/1e6,2), d['parity_check']['ok'])"
}
for t in main sm3 sm5 main sm3 sm5; do run $t III; done
for t in main sm3 sm5; do lib=""; [ "$t" != main ] && lib="QT_LIB_PATH=$PWD/build_ab/$t/libqtesla_b200.so"
  echo "== $t"; env $lib python tools/ab.py --sets III --variants 2 --steps 50 2>&1 | grep -E "cached|ntt_forward|ntt_inverse|natural"; done
