#!/bin/bash
# Build differently configured copies of the CUDA library here (no GPU needed) so that a GPU call only has to run them:
#   tools/variants.sh name1 "-DFOO=1" name2 "-DFOO=2 -DBAR=3" ...   ->  clive2_b200/_lib/variants/<name>.so
# and on the GPU box:  CLV_LIBRARY=clive2_b200/_lib/variants/<name>.so python bench.py ...
cd "$(dirname "$0")/.."
mkdir -p clive2_b200/_lib/variants
FLAGS=$(python -c "from clive2_b200.build import NVCC_FLAGS; print(' '.join(NVCC_FLAGS))")
while [ $# -ge 2 ]; do
  name=$1; extra=$2; shift 2
  ( cd clive2_b200/csrc && nvcc $FLAGS $extra -o ../_lib/variants/$name.so clive_b200.cu > ../_lib/variants/$name.log 2>&1 && echo "built $name [$extra]" || echo "FAILED $name (see _lib/variants/$name.log)" ) &
done
wait
