#!/bin/bash
# A/B of the mixed FP64-quotient / Shoup butterflies (QT_DQ_UNI) in the fused kernels, run r02A
run() { local tag=$1 S=$2; shift 2
  local lib=""; [ "$tag" != main ] && lib="QT_LIB_PATH=$PWD/build_ab/$tag/libqtesla_b200.so"
  env $lib python bench.py --no-extras --set $S --steps 200 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$tag $S', round(d['value']/1e6,2), d['parity_check']['ok'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
}
for t in main dq1 dq2 dq3 dq4 dq2s0 dq1s0; do run $t III; done
for t in main dq1 dq2 dq3; do run $t I; done
