#!/bin/bash
# Builds the format layer with ASan + UBSan and runs the mutation fuzzer on it.
#   tools/fuzz_formats.sh [iterations] [seed files...]
set -e
cd "$(dirname "$0")/.."
mkdir -p build
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer \
    tools/fuzz_formats.cc or-tools_b200/csrc/proto_codec.cc or-tools_b200/csrc/formats.cc or-tools_b200/csrc/params.cc \
    -lz -ldl -o build/fuzz_formats
exec build/fuzz_formats "$@"
