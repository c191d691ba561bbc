#!/bin/bash
# memcheck / racecheck of the kernels WITHOUT a GPU: the emulator build of the library (tests/cuda_emu) compiled with
# AddressSanitizer or ThreadSanitizer, the emulated test-suite run on it.  "Device" buffers are ordinary heap blocks
# here, so an out-of-bounds access of a kernel is a heap-buffer-overflow with file:line of the (rewritten) .cu source.
# usage: tools/emu_sanitize.sh [address|thread] [pytest args...]      -> /tmp/emu_sanitize_<kind>.log
kind=${1:-address}; shift
lib=$(gcc -print-file-name=lib$([ "$kind" = thread ] && echo tsan || echo asan).so)
log=/tmp/emu_sanitize_${kind}.log
# default: the ABI-level emulator tests (the host-layer tests throw C++ exceptions inside torch, which a preloaded
# ASan runtime cannot intercept: "CHECK failed ... real___cxa_throw"); blocks of a launch run one after the other in the
# emulator, so ThreadSanitizer sees races WITHIN a block (shared memory, missing __syncthreads / __syncwarp), not between blocks
if [ $# -eq 0 ]; then set -- tests/test_emu_raster.py tests/test_emu_part.py tests/test_emu_ops.py tests/test_emu_densify.py tests/test_emu_extract.py tests/test_emu_optim.py; fi
PGS_EMU_SANITIZE=$kind LD_PRELOAD=$lib ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 \
  TSAN_OPTIONS=halt_on_error=0:report_signal_unsafe=0:history_size=2 \
  python -m pytest "$@" -q -s -p no:cacheprovider > $log 2>&1
echo "pytest rc=$?  (log: $log)"
echo "sanitizer reports inside the emulated library:"
grep -E "ERROR: AddressSanitizer|WARNING: ThreadSanitizer" $log | sort | uniq -c | head
grep -A12 -E "ERROR: AddressSanitizer|WARNING: ThreadSanitizer" $log | grep -E "pgs::|_build/full" | sort | uniq -c | sort -rn | head -20
tail -3 $log
