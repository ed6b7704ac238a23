#!/bin/bash
# development build of the library into gpurun_out-independent path: tools/devbuild.sh NAME [nvcc flags...]
# -> pyjac_b200/_build/dev_NAME.so (GS = 8 instantiations only); run with PYJAC_B200_LIB=that path
set -e
cd "$(dirname "$0")/.."
name=$1; shift
# phase skipping / clocks need a -DPJ_DEV build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -Xcompiler -O2 \
  ${DEVFULL:--DPJ_DEV_GS8_ONLY} "$@" -I include -I pyjac_b200/csrc -o pyjac_b200/_build/dev_$name.so pyjac_b200/csrc/pyjac_b200.cu
