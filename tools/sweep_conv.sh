#!/bin/bash
# tile-shape sweep of the tcgen05 conv on one layer shape: bash tools/sweep_conv.sh CIN COUT H W
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
for cfg in 256,2,1 256,1,1 256,2,2 256,1,2 128,2,1 128,1,1 128,2,2 128,1,2 64,2,1 64,2,2 64,1,2; do
  bn=${cfg%%,*}; if [ $(( $2 % bn )) -ne 0 ]; then continue; fi
  echo -n "cfg $cfg: "; MAUA_CONV_FORCE=$cfg timeout 60 python tools/prof_conv.py $1 $2 $3 $4 6 2>&1 | tail -1
done
