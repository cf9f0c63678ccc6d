#!/bin/bash
# The CPU oracle (test infrastructure) under ThreadSanitizer and under ASan + UBSan: the checker is multithreaded
# C++ (Sharder-style shard parallelism), so a race or an out-of-bounds read in it would make every parity
# verdict suspect. Builds oracle/pdlp_cpu_solver.cc twice into build/ and runs the oracle-side tests against
# each build through PDLP_ORACLE_LIBRARY.
#   tools/tsan_oracle.sh
set -e
cd "$(dirname "$0")/.."
G=$(dirname "$(gcc -print-file-name=libtsan.so)")
mkdir -p build/tsan build/asan_oracle
TESTS="tests/test_oracle_goldens.py tests/test_kernel_goldens.py tests/test_solver_goldens.py tests/test_feasibility_polishing.py
       tests/test_termination.py tests/test_params_validation.py"
g++ -O1 -g -std=c++17 -fPIC -shared -pthread -fsanitize=thread oracle/pdlp_cpu_solver.cc -o build/tsan/libpdlp_oracle.so
PDLP_ORACLE_LIBRARY=$PWD/build/tsan/libpdlp_oracle.so LD_PRELOAD=$G/libtsan.so TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0" \
  python -m pytest $TESTS -q -m "not gpu" -p no:cacheprovider 2>&1 | tee build/tsan/run.log | tail -2
echo "ThreadSanitizer warnings: $(grep -c 'WARNING: ThreadSanitizer' build/tsan/run.log || true)"
g++ -O1 -g -std=c++17 -fPIC -shared -pthread -fsanitize=address,undefined -fno-omit-frame-pointer oracle/pdlp_cpu_solver.cc -o build/asan_oracle/libpdlp_oracle.so
PDLP_ORACLE_LIBRARY=$PWD/build/asan_oracle/libpdlp_oracle.so LD_PRELOAD=$G/libasan.so:$G/libubsan.so ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
  UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=0 python -m pytest $TESTS tests/test_synthetic_configs.py -q -m "not gpu" -p no:cacheprovider 2>&1 | tee build/asan_oracle/run.log | tail -2
echo "ASan / UBSan reports: $(grep -c 'runtime error\|ERROR: AddressSanitizer' build/asan_oracle/run.log || true)"
