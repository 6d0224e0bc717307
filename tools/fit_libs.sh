#!/bin/bash
# A/B of fit-kernel builds: every libpmk_b200*.so in the package directory, fit timings on the workloads given (default c3)
# usage: tools/fit_libs.sh [workload ...]     (environment, e.g. PMK_CHOL_VARIANT=0, is passed through)
wl=${@:-c3}
for L in "" $(ls patchmixturekriging_b200/libpmk_b200_*.so 2>/dev/null); do
  for w in $wl; do
    echo -n "${L:-product} $w: "; PMK_LIB=${L:+$PWD/$L} timeout 300 python tools/fit_only.py $w 3 2>&1 | tail -1
  done
done
