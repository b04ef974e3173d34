#!/bin/bash
# Builds the library AND its phase-timing twin (gp_fit.cu with -DTLBO_PHASE_TIMING, the other objects reused), in parallel:
#   tools/build_timing.sh && TLBO_B200_LIB=efficient-tlbo_b200/lib/libtlbo_b200_timing.so python tools/phase_fit.py 60
set -e
cd "$(dirname "$(readlink -f "$0")")/../efficient-tlbo_b200"
mkdir -p build_timing
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -ffp-contract=off \
  -diag-suppress 550 -DTLBO_PHASE_TIMING -c csrc/gp_fit.cu -o build_timing/gp_fit.o 2>&1 | grep -v deprecated &
python -m tlbo_b200.build >/dev/null
wait
nvcc -shared -o lib/libtlbo_b200_timing.so build_timing/gp_fit.o build/c_api.o build/gp_mcmc.o build/gp_predict.o \
  build/host_search.o build/weights.o 2>&1 | grep -v deprecated || true
ls -la lib/
