#!/bin/bash
# The host side of libpdlp_b200.so under AddressSanitizer + UBSan, without a GPU:
# every .cc of csrc/ is rebuilt with -fsanitize=address,undefined, linked with the device objects of the
# normal build into build/asan/libpdlp_b200.so, and the CPU tests that call the library's host-only entry
# points (validation, termination predicates, SELL layout, row blocks, peer-arena layout, format layer)
# run against it through PDLP_B200_LIBRARY. The device code itself is compute-sanitizer's job (tools/sanitize.sh).
#   tools/asan_host.sh [pytest args...]
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > /dev/null
OUT=build/asan
mkdir -p $OUT
GCC_LIB=$(dirname "$(gcc -print-file-name=libasan.so)")
for f in sell_builder device_problem solver params capi comm proto_codec formats; do
  g++ -std=c++17 -O1 -g -fPIC -pthread -fsanitize=address,undefined -fno-omit-frame-pointer -I/usr/local/cuda/include \
      -c or-tools_b200/csrc/$f.cc -o $OUT/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libpdlp_b200.so build/obj/device_ops.cu.o build/obj/device_build.cu.o \
    $OUT/sell_builder.o $OUT/device_problem.o $OUT/solver.o $OUT/params.o $OUT/capi.o $OUT/comm.o $OUT/proto_codec.o $OUT/formats.o \
    -ldl -lpthread -lz -L"$GCC_LIB" -Xlinker -lasan -Xlinker -lubsan
TESTS="tests/test_boundary.py tests/test_termination.py tests/test_params_validation.py tests/test_sell_layout.py tests/test_native_io.py
       tests/test_distributed.py tests/test_host_logic_properties.py tests/test_python_surface.py tests/test_problem_io.py"
# (leak detection off: the interpreter itself never frees everything; test_cpp_example links its own binaries against the normal build)
PDLP_B200_LIBRARY=$PWD/$OUT/libpdlp_b200.so LD_PRELOAD="$GCC_LIB/libasan.so:$GCC_LIB/libubsan.so" \
  ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=0 \
  python -m pytest $TESTS -q -m "not gpu" -p no:cacheprovider "$@" 2>&1 | tee $OUT/run.log | tail -4
echo "sanitizer reports: $(grep -c 'runtime error\|ERROR: AddressSanitizer' $OUT/run.log || true)"
