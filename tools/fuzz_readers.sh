#!/bin/bash
# tools/fuzz_readers.py against the ASan + UBSan build of the host layer (tools/asan_host.sh builds it).
# Usage: bash tools/fuzz_readers.sh [iterations per file] [seed]
cd "$(dirname "$0")/.."
D=sassena_b200/build/asan
[ -f $D/libsassena_b200_asan.so ] || { echo "run tools/asan_host.sh first"; exit 2; }
mkdir -p gpurun_out; rm -f gpurun_out/asan_report.* gpurun_out/ubsan_report.*
ASAN_OPTIONS=detect_leaks=0:new_delete_type_mismatch=0:alloc_dealloc_mismatch=0:allocator_may_return_null=1:log_path=$PWD/gpurun_out/asan_report \
UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=0:log_path=$PWD/gpurun_out/ubsan_report \
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so.6)" \
SASSENA_B200_LIB=$PWD/$D/libsassena_b200_asan.so timeout 3000 python tools/fuzz_readers.py "${1:-300}" "${2:-1}"
echo "exit $?"
ls gpurun_out | grep -E "asan_report|ubsan_report" || echo "no sanitizer reports"
