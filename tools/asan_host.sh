#!/bin/bash
# The C++ host layer (csrc/host/*.cpp: control plane, readers, HDF5 writer, decomposition, devices) under AddressSanitizer +
# UndefinedBehaviorSanitizer: builds a second library next to the product (kernel objects reused from sassena_b200/build/) and
# runs the CPU test suite against it.  The two mismatch checks are off because the reference's own code in oracle/_ref trips
# them (ref_select_pdb deletes through a base class without a virtual destructor).
set -e
cd "$(dirname "$0")/.."
python -m sassena_b200.build > /dev/null
D=sassena_b200/build/asan; mkdir -p $D
for f in sassena_b200/csrc/host/*.cpp; do
  /usr/bin/g++ -std=c++17 -O1 -g -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -I include -I sassena_b200/csrc \
    -I /usr/local/cuda/include -c $f -o $D/$(basename $f).o &
done; wait
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a --shared -o $D/libsassena_b200_asan.so \
  sassena_b200/build/kernels_*.o sassena_b200/build/nccl_api.cpp.o sassena_b200/build/sgpu_capi.cu.o $D/*.o -ldl -Xlinker -lasan -Xlinker -lubsan
mkdir -p gpurun_out; rm -f gpurun_out/asan_report.* gpurun_out/ubsan_report.*
ASAN_OPTIONS=detect_leaks=0:new_delete_type_mismatch=0:alloc_dealloc_mismatch=0:log_path=$PWD/gpurun_out/asan_report \
UBSAN_OPTIONS=print_stacktrace=1:log_path=$PWD/gpurun_out/ubsan_report \
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so.6)" \
SASSENA_B200_LIB=$PWD/$D/libsassena_b200_asan.so python -m pytest tests -q -m "not gpu" -p no:cacheprovider
ls gpurun_out | grep -E "asan_report|ubsan_report" || echo "no sanitizer reports"
