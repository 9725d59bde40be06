#!/bin/bash
# profiles/build_variant.sh NAME [-DIMFEAT_X=.. ...] -- a build of the library with other compile-time switches, for
# profiles/sweep_run.py ("LIB=profiles/_variants/NAME.so" in an --env spec); development only.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p profiles/_variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -Xptxas -v "$@" \
  -o profiles/_variants/$name.so interpretable-multichannel-image-analysis_b200/csrc/imfeat_api.cu 2> profiles/_variants/$name.ptxas.txt
grep -A2 "k4w_shape_kernelILb1\|k12_basic_kernelILb1" profiles/_variants/$name.ptxas.txt | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | paste - - - -
