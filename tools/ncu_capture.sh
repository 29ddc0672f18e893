#!/bin/bash
# ncu --set full captures of the hot kernels, one .ncu-rep + one JSON summary each (run under gpurun, ONE GPU).
# Usage: bash tools/ncu_capture.sh <tag> [which ...]      which: fused fusedI fusedpI fusedpIII nussF64 nussRing nussRec nussP3 nussP3ring
# The programs are run once WITHOUT ncu first (a number printed under ncu is never a bench value).
TAG=${1:-r02}; shift
WHICH=${@:-fused fusedpI fusedpIII}
OUT=gpurun_out; mkdir -p $OUT
cap() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  "$@" > $OUT/ncu_plain_${name}_$TAG.log 2>&1 || { echo "$name: plain run failed"; tail -3 $OUT/ncu_plain_${name}_$TAG.log; return; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o $OUT/prof_${name}_$TAG "$@" > $OUT/ncu_${name}_$TAG.log 2>&1
  echo "$name: ncu rc=$?"
  ncu -i $OUT/prof_${name}_$TAG.ncu-rep --page source --csv > $OUT/ncu_${name}_${TAG}_source.csv 2>/dev/null
  python tools/ncu_summary.py $OUT/prof_${name}_$TAG.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 ($*), run $TAG" > $OUT/ncu_${name}_$TAG.json 2>> $OUT/ncu_${name}_$TAG.log
  python - <<PY
import json
d=json.load(open("$OUT/ncu_${name}_$TAG.json")); m=d["metrics"]
g=lambda k: m.get(k,{}).get("value")
print("  ", d["kernel"][:60], "us", g("gpu__time_duration.sum"), "regs", g("launch__registers_per_thread"), "fmaheavy%", g("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
      "alu%", g("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"), "fp64%", g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), "issue%", g("smsp__issue_active.avg.pct_of_peak_sustained_active"))
print("   stalls/issue:", {k: round(v,2) for k,v in sorted(d["stall_per_issue"].items(), key=lambda kv:-kv[1])[:7]})
PY
  [ -n "$KEEP_REP" ] || rm -f $OUT/prof_${name}_$TAG.ncu-rep   # gpurun copies back at most 64 MiB: keep the CSV pages, not the 27 MB report
}
for w in $WHICH; do
  case $w in
    fused)      cap fused      k_polymul_tma   4 python bench.py --steps 5 --warmup 3 --no-extras ;;
    fusedI)     cap fusedI     k_polymul_tma   4 python bench.py --steps 5 --warmup 3 --no-extras --set I ;;
    fusedpI)    cap fusedpI    k_polymul_tma   4 python bench.py --steps 5 --warmup 3 --no-extras --set p-I ;;
    fusedpIII)  cap fusedpIII  k_polymul_pair 4 python bench.py --steps 5 --warmup 3 --no-extras --set p-III ;;
    nussF64)    cap nussF64    k_nussbaumer_warp 2 python tools/nuss_one.py III 1 3 ;;
    nussRing)   cap nussRing   k_nussbaumer_warp 2 python tools/nuss_one.py III 0 0 ;;
    nussRec)    cap nussRec    k_nussbaumer_warp 2 python tools/nuss_one.py III 1 2 ;;
    nussSchool) cap nussSchool k_nussbaumer_warp 2 python tools/nuss_one.py III 1 1 ;;
    nussP3)     cap nussP3     k_nussbaumer 2 python tools/nuss_one.py p-III 1 0 ;;
    nussP3ring) cap nussP3ring k_nussbaumer 2 python tools/nuss_one.py p-III 0 0 ;;
    nussPI)     cap nussPI     k_nussbaumer_warp 2 python tools/nuss_one.py p-I 1 0 ;;
  esac
done
ls -la $OUT/*.ncu-rep 2>/dev/null | tail -12
