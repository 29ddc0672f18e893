#!/bin/bash
# A/B of the uniform warp-index idiom (QT_UNIFORM_WARP / QT_NUSS_UNIFORM_WARP), run r02F: main = on, nouni = off
run() { local tag=$1 S=$2; shift 2
  local lib=""; [ "$tag" != main ] && lib="QT_LIB_PATH=$PWD/build_ab/$tag/libqtesla_b200.so"
  env $lib python bench.py --no-extras --set $S --steps 200 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$tag $S', round(d['value']/1e6,2), d['parity_check']['ok'])"
}
for S in III I p-I p-III; do for t in main nouni main nouni; do run $t $S; done; done
for cfg in "III 1 3" "III 0 0" "I 1 3" "p-I 1 0" "p-III 1 0"; do for t in main nouni; do
  lib=""; [ "$t" != main ] && lib="QT_LIB_PATH=$PWD/build_ab/$t/libqtesla_b200.so"
  echo -n "$t: "; env $lib python tools/nuss_one.py $cfg; done; done
for t in main nouni; do lib=""; [ "$t" != main ] && lib="QT_LIB_PATH=$PWD/build_ab/$t/libqtesla_b200.so"
  echo "== $t unfused"; env $lib python tools/ab.py --sets III,p-III --variants 2 --steps 50 2>&1 | grep -E "cached|ntt_forward|ntt_inverse|natural|variant"; done
