#!/bin/bash
# warps per CTA of the fused kernels re-measured after the uniform warp-index change (fewer registers), run r02G
run() { local tag=$1 S=$2; shift 2
  local lib=""; [ "$tag" != main ] && lib="QT_LIB_PATH=$PWD/build_ab/$tag/libqtesla_b200.so"
  env $lib python bench.py --no-extras --set $S --steps 200 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$tag $S', round(d['value']/1e6,2), d['parity_check']['ok'])"
}
for t in main w20I main w20I; do run $t I; done
for t in main w24III w16III main w24III; do run $t III; done
