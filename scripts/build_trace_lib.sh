#!/bin/sh
# Instrumented build of the library (CTA-0 timeline of conv_igemm_kernel) into build/trace/; use with
#   CSBSR_LIB_PATH=build/trace/libcsbsr_b200.so python scripts/trace_conv.py <case>
set -e
cd "$(dirname "$0")/.."
mkdir -p build/trace
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
for f in csbsr_b200/csrc/*.cu; do
  o=build/trace/$(basename "$f" .cu).o
  if [ "$(basename "$f")" = conv_igemm.cu ]; then
    nvcc $FLAGS -DCSBSR_CONV_TRACE_BUILD -c "$f" -o "$o"
  elif [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then
    nvcc $FLAGS -c "$f" -o "$o"
  fi
done
nvcc -shared -o build/trace/libcsbsr_b200.so build/trace/*.o -lcudart
