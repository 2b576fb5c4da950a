#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Builds the HOST layer of libperseus_gpu (handle.cu, stream_path.cu, bulk_path.cu, perseus_vrx.cpp, perseus_host.cpp,
# copy_pool.cpp) against the CUDA stand-in of tests/sanitize/fake_cuda into a shared library with the product's C ABI, so the slab
# ring, the delivery thread, eager submission, the watchdog and the staging pipeline can be driven from Python on a box without a
# GPU (tests/test_hostsim_cpu.py).  The "kernels" of this build call the CPU oracle; the product never links any of this.
#   tests/hostsim/build.sh <output .so>
set -eu
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
OUT="$1"
CSRC="$ROOT/libperseus-sdr_b200/csrc"
CXX=/usr/bin/g++
[ -x "$CXX" ] || CXX=g++
TMP="$(mktemp -d)"
gcc -O2 -fPIC -c "$ROOT/oracle/perseus_oracle.c" -o "$TMP/oracle.o"
$CXX -std=c++17 -O2 -fPIC -shared -pthread -I"$ROOT/tests/sanitize/fake_cuda" -I"$CSRC" \
	-x c++ "$CSRC/handle.cu" "$CSRC/stream_path.cu" "$CSRC/bulk_path.cu" -x none "$ROOT/tests/sanitize/fake_cuda.cpp" "$CSRC/perseus_vrx.cpp" "$CSRC/perseus_host.cpp" "$CSRC/copy_pool.cpp" \
	"$TMP/oracle.o" -o "$OUT"
rm -rf "$TMP"
