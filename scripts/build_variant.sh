#!/bin/bash
# build_variant.sh <name> [git-rev]: libomr_b200.so of a git revision (default HEAD) into build_variants/<name>/lib.so, for A/B runs
# with OMR_B200_LIB (scripts/ab_stage.sh).  build_variants/ is git-ignored but travels to the GPU box.
set -e
NAME=$1; REV=${2:-HEAD}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
D=$ROOT/build_variants/$NAME
rm -rf $D; mkdir -p $D/tfhe-omr_b200 $D/include
git -C $ROOT archive $REV tfhe-omr_b200/csrc include | tar -x -C $D
CUT=$(python -c "
import importlib.util
spec=importlib.util.spec_from_file_location('b','$ROOT/tfhe-omr_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); print(m._cutlass_include())")
cd $D/tfhe-omr_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -cudart static -DOMR_HAVE_CUTLASS --expt-relaxed-constexpr \
  -diag-suppress 20012 -I$CUT/include -I$CUT/tools/util/include -o $D/lib.so $(ls *.cu) -ldl
ls -la $D/lib.so
