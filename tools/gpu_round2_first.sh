#!/bin/bash
# First GPU call of round 2: the tuning builds prepared (but not measured / only spot-measured) at the end of round 1.
#   HERE:  bash tools/gpu_round2_first.sh build        # makes rails_b200/lib/libmol_b200_<name>.so for every variant below
#   then:  gpurun --timeout 900 -- 'bash tools/gpu_round2_first.sh run'
# `run` = (1) tools/quick_variants.py: every build in one process - ms per 512 x 1M step, final ids / scores against the
# default build, coarse-score difference; a hanging variant is cut by its own timeout and the rest re-run without it;
# (2) the GPU parity tests that touch the coarse pass, for the two fastest candidates (edit CANDIDATES after reading (1)).
VARIANTS=(
  "h2_3e MOL_E2_H2_MASK=0x3E MOL_E2_POLY_MASK=0"       # measured in round 1: 34.37 vs 35.37 ms
  "h2_1e MOL_E2_H2_MASK=0x1E MOL_E2_POLY_MASK=0"
  "h2_3c MOL_E2_H2_MASK=0x3C MOL_E2_POLY_MASK=0"
  "h2_7e MOL_E2_H2_MASK=0x7E MOL_E2_POLY_MASK=0"
  "h2_3f MOL_E2_H2_MASK=0x3F MOL_E2_POLY_MASK=0"
  "g1late MOL_G1_LATE=1"                                # next query's G1 issued behind G3a instead of behind G2 (unmeasured)
  "g1late_h2_3e MOL_G1_LATE=1 MOL_E2_H2_MASK=0x3E MOL_E2_POLY_MASK=0"
  "sh1 MOL_E2_SHARE=1"                                  # E3 group converts E2's last chunk (unmeasured)
  "sh2 MOL_E2_SHARE=2"
  "sh1_h2_3e MOL_E2_SHARE=1 MOL_E2_H2_MASK=0x3E MOL_E2_POLY_MASK=0"
  "sh1_h2_be MOL_E2_SHARE=1 MOL_E2_H2_MASK=0xBE MOL_E2_POLY_MASK=0"  # the shared chunk (7) MUFU-free as well
  "all3 MOL_G1_LATE=1 MOL_E2_SHARE=1 MOL_E2_H2_MASK=0x3E MOL_E2_POLY_MASK=0"
)
CANDIDATES=${CANDIDATES:-"h2_3e sh1_h2_3e"}
cd "$(dirname "$0")/.."
if [ "$1" = "build" ]; then
  for v in "${VARIANTS[@]}"; do set -- $v; python -m rails_b200.build --variant "$@" | tail -1; done
  exit 0
fi
mkdir -p gpurun_out
names="default"; for v in "${VARIANTS[@]}"; do set -- $v; names="$names $1"; done
# one process for all; experimental synchronisation (sh*) goes last so that a hang costs only those
timeout 240 python tools/quick_variants.py $names 2>&1 | tail -20
for v in $CANDIDATES; do
  MOL_B200_LIB=$PWD/rails_b200/lib/libmol_b200_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_$v.log 2>&1
  echo "$v pytest exit $?"; tail -3 gpurun_out/pytest_$v.log
done
