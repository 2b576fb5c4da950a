#!/bin/bash
# Host-layer sanitizer runs: builds the C-ABI host code -- handle.cu / stream_path.cu / bulk_path.cu (no kernels in them), perseus_vrx.cpp,
# perseus_host.cpp -- with a plain C++ compiler against the CUDA stand-in in tests/sanitize/fake_cuda, once with
# ThreadSanitizer and once with AddressSanitizer + UBSan, and runs tests/sanitize/host_stress.cpp under each.
# No GPU involved; the device kernels have their own compute-sanitizer runs (profiles/sanitizer_*).
#   tests/sanitize/sanitize.sh [outdir]      logs -> <outdir>/r2_sanitizer_host_{tsan,asan}.txt   (default: profiles/)
set -u
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
OUT="${1:-$ROOT/profiles}"
BUILD="$(mktemp -d)"
CSRC="$ROOT/libperseus-sdr_b200/csrc"
CXX=/usr/bin/g++        # the image exports CXX=/opt/gcc/bin/g++, a wrapper without the sanitizer runtimes
SRCS="$ROOT/tests/sanitize/host_stress.cpp $ROOT/tests/sanitize/fake_cuda.cpp $CSRC/perseus_vrx.cpp $CSRC/perseus_host.cpp $CSRC/copy_pool.cpp"
rc=0
for san in tsan asan; do
	case $san in
	tsan) FLAGS="-fsanitize=thread" ;;
	asan) FLAGS="-fsanitize=address,undefined -fno-sanitize-recover=undefined" ;;
	esac
	log="$OUT/r2_sanitizer_host_$san.txt"
	{
		echo "# $san: $CXX -O1 -g $FLAGS  (host layer of libperseus_gpu over tests/sanitize/fake_cuda; $(date -u +%FT%TZ))"
		/usr/bin/gcc -O1 -g $FLAGS -fPIC -c "$ROOT/oracle/perseus_oracle.c" -o "$BUILD/oracle_$san.o" &&
		$CXX -std=c++17 -O1 -g $FLAGS -pthread -Wall -Wextra -Wno-tsan -I"$ROOT/tests/sanitize/fake_cuda" -I"$CSRC" \
			-x c++ "$CSRC/handle.cu" "$CSRC/stream_path.cu" "$CSRC/bulk_path.cu" -x none $SRCS "$BUILD/oracle_$san.o" -o "$BUILD/host_stress_$san" &&
		for nomb in 0 1; do   # the callback / owner hand-off has two implementations: sys_membarrier (asymmetric) and plain fences
			echo "# PERSEUS_GPU_NO_MEMBARRIER=$nomb PERSEUS_GPU_NO_TSC=$nomb PERSEUS_GPU_NT_COPY=$([ $nomb = 1 ] && echo sse2 || echo widest)"
			# ... and the callback's clock and the slab copy two each: TSC-carried / clock_gettime, widest / 16-byte stores
			PERSEUS_GPU_NO_TSC=$nomb PERSEUS_GPU_NT_COPY=$([ $nomb = 1 ] && echo sse2) PERSEUS_GPU_NO_MEMBARRIER=$nomb TSAN_OPTIONS="halt_on_error=0 second_deadlock_stack=1" ASAN_OPTIONS="detect_leaks=1" "$BUILD/host_stress_$san"
			status=$?
			echo "# exit status $status"
			[ $status -eq 0 ] || rc=1
		done
	} > "$log" 2>&1
	grep -q "WARNING: ThreadSanitizer\|ERROR: AddressSanitizer\|runtime error\|FAILED\|LeakSanitizer" "$log" && rc=1
	tail -3 "$log"
done
rm -rf "$BUILD"
exit $rc
