#!/bin/bash
# tools/slab_sweep.sh: virtual-slab runs over (n, ranks) and kernel toggles, each in its own process (a faulting
# kernel poisons the CUDA context); prints one status line per case.
run() { # label, env..., n, ranks
  local label=$1; shift
  local out
  out=$(env "${@:1:$#-2}" timeout 120 python tools/virtual_slab_run.py "${@: -2:1}" "${@: -1}" 1 2>&1 | tail -1 | cut -c1-150)
  echo "$label n=${@: -2:1} ranks=${@: -1}: $out"
}
for c in "5793 8" "11585 4" "11585 8" "11585 7" "11585 6" "8193 4" "5793 2" "9001 5" "7001 3"; do
  run plain X=1 $c
done
for t in NF_RBSOR_TMA=0; do
  run $t $t 11585 8
done
