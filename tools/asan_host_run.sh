#!/bin/bash
# Host-side AddressSanitizer + UBSan run of the library's C++ (capi.cu / capi_multi.cuh host code), the C++ host mirror with
# its tests, and the plain-C example.  compute-sanitizer is closed on the GPU pool; this covers the host half: handle
# lifetimes, the staging/plan/offset vectors, the group's worker threads.  Build here (no GPU needed), run on a GPU box:
#   tools/asan_host_run.sh build      then      gpurun -- tools/asan_host_run.sh run
set -e
cd "$(dirname "$0")/.."
D=build/asan
if [ "$1" = build ]; then
  mkdir -p $D
  (cd codex-storage-proofs-circuits_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -lineinfo -std=c++17 \
     -Xcompiler -fPIC,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer -shared -ldl -o ../../$D/libcodexcommit.so capi.cu)
  g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -o $D/test_host tests/host_cpp/test_host.cpp \
     codex-storage-proofs-circuits_b200/host/proof_input.cpp -L$D -lcodexcommit -lpthread -Wl,-rpath,'$ORIGIN'
  gcc -std=c99 -O1 -g -fsanitize=address,undefined -Iinclude examples/commit_dataset.c -L$D -lcodexcommit -Wl,-rpath,'$ORIGIN' -o $D/commit_dataset
  exit 0
fi
export ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1
mkdir -p gpurun_out
$D/commit_dataset > gpurun_out/asan_example.log 2>&1; echo "example rc=$?"
$D/test_host > gpurun_out/asan_test_host.log 2>&1; echo "test_host rc=$?"
echo "sanitizer reports: $(cat gpurun_out/asan_example.log gpurun_out/asan_test_host.log | grep -c 'ERROR: AddressSanitizer\|runtime error' || true)"
tail -2 gpurun_out/asan_test_host.log
