#!/bin/bash
# time the pair kernel of every experiment build present (PMK_VARIANT builds of build.py) on a workload
W=${1:-c3_mini}
timeout 60 python tools/k3_time.py $W
for f in patchmixturekriging_b200/libpmk_b200_*.so; do
  case $f in *prof*) continue;; esac
  PMK_LIB=$PWD/$f timeout 60 python tools/k3_time.py $W
done
