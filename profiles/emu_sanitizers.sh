#!/bin/bash
# compute-sanitizer is closed on the GPU pool ("closed on this pool and stays closed", gpurun_out/r2a_sanitizer_*.log), so the race /
# bounds evidence comes from the CPU emulation of the same kernel sources (tests/emu/cuda_emu.h: one pthread per CUDA thread,
# __syncthreads / __syncwarp / warp collectives = pthread barriers) built with ThreadSanitizer and with AddressSanitizer:
#   TSan  = racecheck + synccheck of the kernel logic (two threads touching the same shared/global word without a barrier between them;
#           it is stricter than the hardware: lanes of a warp are NOT lock-stepped here, so every missing __syncwarp shows too)
#   ASan  = memcheck (out-of-bounds shared / global accesses of the emulated kernels)
# Not covered: the tcgen05 / TMA / mbarrier instructions themselves (no CPU emulation); their protocol is bounded by trapping waits
# (fpt_mbar_wait) and pinned by the 1000-permutation parity tests against the oracle.
cd "$(dirname "$0")/.."
OUT=profiles
TSAN_SO=$(gcc -print-file-name=libtsan.so); ASAN_SO=$(gcc -print-file-name=libasan.so)
sh tests/emu/build.sh "$PWD/tests/emu/libfpt_emu_tsan.so" -fsanitize=thread -O1 || exit 1
sh tests/emu/build.sh "$PWD/tests/emu/libfpt_emu_asan.so" -fsanitize=address -O1 || exit 1
sh tests/emu/build.sh "$PWD/tests/emu/libfpt_emu_tsan_exact.so" -fsanitize=thread -O1 -DFPT_SMACOF_BOUND_SCALE=1e12 || exit 1
sh tests/emu/build.sh "$PWD/tests/emu/libfpt_emu_asan_exact.so" -fsanitize=address -O1 -DFPT_SMACOF_BOUND_SCALE=1e12 || exit 1
SEL="${1:-}"
# the detectors must be live: a deliberately racy / overrunning toy kernel in the same library has to be reported
TSAN_OPTIONS="halt_on_error=0 exitcode=0 log_path=/tmp/fpt_tsan_self" LD_PRELOAD=$TSAN_SO python -c "
import ctypes, numpy as np
l = ctypes.CDLL('$PWD/tests/emu/libfpt_emu_tsan.so'); o = np.zeros(64, dtype=np.int32)
l.emu_selftest(1, 0, o.ctypes.data_as(ctypes.c_void_p))" ; CLEAN=$(cat /tmp/fpt_tsan_self.* 2>/dev/null | grep -c 'WARNING: ThreadSanitizer'); rm -f /tmp/fpt_tsan_self.*
TSAN_OPTIONS="halt_on_error=0 exitcode=0 log_path=/tmp/fpt_tsan_self" LD_PRELOAD=$TSAN_SO python -c "
import ctypes, numpy as np
l = ctypes.CDLL('$PWD/tests/emu/libfpt_emu_tsan.so'); o = np.zeros(64, dtype=np.int32)
l.emu_selftest(0, 0, o.ctypes.data_as(ctypes.c_void_p))" ; RACY=$(cat /tmp/fpt_tsan_self.* 2>/dev/null | grep -c 'WARNING: ThreadSanitizer'); rm -f /tmp/fpt_tsan_self.*
ASAN_OPTIONS="detect_leaks=0 halt_on_error=0 log_path=/tmp/fpt_asan_self" LD_PRELOAD=$ASAN_SO python -c "
import ctypes, numpy as np
l = ctypes.CDLL('$PWD/tests/emu/libfpt_emu_asan.so'); o = np.zeros(64, dtype=np.int32)
l.emu_selftest(1, 1, o.ctypes.data_as(ctypes.c_void_p))" ; OVER=$(cat /tmp/fpt_asan_self.* 2>/dev/null | grep -c 'ERROR: AddressSanitizer'); rm -f /tmp/fpt_asan_self.*
echo "detector self-test: TSan reports with barrier = $CLEAN (want 0), without barrier = $RACY (want > 0); ASan reports on a deliberate overrun = $OVER (want > 0)" | tee $OUT/r2_emu_sanitizer_selftest.log
TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0 log_path=/tmp/fpt_tsan" LD_PRELOAD=$TSAN_SO FPT_EMU_LIB=$PWD/tests/emu/libfpt_emu_tsan.so FPT_EMU_LIB_EXACT=$PWD/tests/emu/libfpt_emu_tsan_exact.so \
  timeout 7200 python -m pytest tests/test_emu_kernels.py -q -x -p no:cacheprovider ${SEL:+-k "$SEL"} > $OUT/r2_emu_tsan_pytest.log 2>&1
cat /tmp/fpt_tsan.* > $OUT/r2_emu_tsan_reports.log 2>/dev/null; rm -f /tmp/fpt_tsan.*
echo "ThreadSanitizer reports: $(grep -c 'WARNING: ThreadSanitizer' $OUT/r2_emu_tsan_reports.log 2>/dev/null)" >> $OUT/r2_emu_tsan_pytest.log
ASAN_OPTIONS="detect_leaks=0 halt_on_error=0 log_path=/tmp/fpt_asan" LD_PRELOAD=$ASAN_SO FPT_EMU_LIB=$PWD/tests/emu/libfpt_emu_asan.so FPT_EMU_LIB_EXACT=$PWD/tests/emu/libfpt_emu_asan_exact.so \
  timeout 7200 python -m pytest tests/test_emu_kernels.py -q -x -p no:cacheprovider ${SEL:+-k "$SEL"} > $OUT/r2_emu_asan_pytest.log 2>&1
cat /tmp/fpt_asan.* > $OUT/r2_emu_asan_reports.log 2>/dev/null; rm -f /tmp/fpt_asan.*
echo "AddressSanitizer reports: $(grep -c 'ERROR: AddressSanitizer' $OUT/r2_emu_asan_reports.log 2>/dev/null)" >> $OUT/r2_emu_asan_pytest.log
tail -n 3 $OUT/r2_emu_tsan_pytest.log; tail -n 3 $OUT/r2_emu_asan_pytest.log
