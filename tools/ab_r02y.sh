#!/bin/bash
run() { local tag=$1 S=$2; shift 2
  local lib=""; [ "$tag" != main ] && lib="QT_LIB_PATH=$PWD/build_ab/$tag/libqtesla_b200.so"
  env $lib python bench.py --no-extras --set $S --steps 200 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$tag $S', round(d['value']/1e6,2), d['parity_check']['ok'])"
}
for t in main pair6 pair10 pair12; do run $t p-III; done
