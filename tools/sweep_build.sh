#!/bin/bash
# usage: tools/sweep_build.sh "-DFOO=1 -DBAR=2" ...  -> rebuild with each flag set and print one bench line
for flags in "$@"; do
  CLV_NVCC_EXTRA="$flags" python -m clive2_b200.build --force > /dev/null 2>&1 || { echo "$flags BUILD FAILED"; continue; }
  tools/sweep.sh "X=[$(echo $flags | tr ' ' ',')]"
done
python -m clive2_b200.build --force > /dev/null 2>&1
