#!/bin/bash
# Builds the compile-time variants that were prepared (and checked on the CPU) at the end of round 1 but have not run on a GPU yet,
# then -- when a GPU is present -- runs the parity tests and the stage timing on each.  Usage (through gpurun, after the build here):
#   bash tools/round2_candidates.sh build        (no GPU needed)
#   gpurun -- 'bash tools/round2_candidates.sh run'
set -e
cd "$(dirname "$0")/.."
case "$1" in
build)
  bash tools/build_variant.sh ff -DB2F_RESOLVE_FREE_FIRST=1        # k_spec_resolve: matches that read only bytes from before the step are copied unordered
  bash tools/build_variant.sh ck16 -DB2F_CHECKSUM_VEC16=1          # k_checksum: 16-byte loads
  bash tools/build_variant.sh ffck -DB2F_RESOLVE_FREE_FIRST=1 -DB2F_CHECKSUM_VEC16=1
  ;;
run)
  mkdir -p gpurun_out
  for v in "" _ff _ck16 _ffck; do
    echo "== variant '$v'"
    B2F_LIB=libflate_b200/libb2f$v.so timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
    B2F_LIB=libflate_b200/libb2f$v.so timeout -s KILL 120 python tools/stage_times.py 265 A 2>&1 | grep "stages (ms)" | tail -2
  done
  ;;
*) echo "usage: $0 build|run"; exit 1;;
esac
