#!/bin/bash
# The KERNEL SOURCE (curvine_b200/csrc/kernels.cu) under sanitizers, without a GPU: the kernel parity suites run against the SIMT shim
# of tests/simt_emu (a fiber per CUDA thread), built
#   1. with -fsanitize=address,undefined: out-of-bounds stores, loads of granules that lie entirely outside every allocation (the
#      kernels read whole 16-byte aligned granules: simt_emu.h states the rule), misaligned accesses, shift/overflow UB;
#   2. with -fsanitize=thread as a race check: every CUDA thread of the first two warps and every other warp is a TSan fiber, ordered
#      only by what the programming model orders (block start/end, __syncthreads, warp collectives, atomics).
# The shim's own known-answer programs run first in both builds: they must REPORT a planted out-of-bounds store / load and planted
# races (missing __syncthreads, missing __syncwarp-less neighbour exchange, plain add from two blocks) and stay silent on the correct twins.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"; cd "$ROOT"
SUITES="tests/test_kernels_gpu.py tests/test_zzz_new_kernels_gpu.py"
ARCH="$(uname -m)"

E=$(python tests/simt_emu/build.py --selftest address)
o1=$(ASAN_OPTIONS=detect_stack_use_after_return=0:detect_leaks=0 $E oob 1 2>&1 | grep -c "AddressSanitizer: heap-buffer-overflow" || true)
o2=$(ASAN_OPTIONS=detect_stack_use_after_return=0:detect_leaks=0 $E oob 2 2>&1 | grep -c "AddressSanitizer: heap-buffer-overflow" || true)
o3=$(ASAN_OPTIONS=detect_stack_use_after_return=0:detect_leaks=0 $E oob 3 2>&1 | grep -c "AddressSanitizer" || true)
echo "asan selftest: planted OOB store reported=$o1 planted OOB granule load reported=$o2 in-granule over-read reported=$o3 (want >0 >0 0)"
LIB=$(python tests/simt_emu/build.py address,undefined)
CV_TEST_MOCK_CUDA_LIB=$LIB CV_SIMT_EMU_THREADS=8 CV_SIMT_EMU_SMS=16 LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
  ASAN_OPTIONS=detect_leaks=0:halt_on_error=0:detect_stack_use_after_return=0 UBSAN_OPTIONS=print_stacktrace=1 \
  timeout 1800 python -m pytest $SUITES -m gpu -q -p no:cacheprovider > /tmp/cv_kernels_asan.log 2>&1 || true
echo "asan+ubsan: $(tail -1 /tmp/cv_kernels_asan.log) | findings: $(grep -ciE 'runtime error|ERROR: AddressSanitizer' /tmp/cv_kernels_asan.log)"

E=$(python tests/simt_emu/build.py --selftest thread)
r=""
for k in 0 1 2 3 4 5 6; do
  r="$r $(TSAN_OPTIONS='halt_on_error=0 history_size=4' CV_SIMT_EMU_THREADS=2 setarch $ARCH -R $E race $k 2>&1 | grep -c 'ThreadSanitizer: data race' || true)"
done
echo "tsan selftest: races reported for [no barrier, __syncthreads, __syncwarp, two blocks plain add, two blocks atomicAdd, block sum without barrier, with barrier] =$r (want >0 0 0 >0 0 >0 0)"
LIB=$(python tests/simt_emu/build.py thread)
CV_TEST_MOCK_CUDA_LIB=$LIB CV_SIMT_EMU_THREADS=2 CV_SIMT_EMU_SMS=6 LD_PRELOAD="$(gcc -print-file-name=libtsan.so)" TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4" \
  timeout 2400 setarch $ARCH -R python -m pytest $SUITES -m gpu -q -p no:cacheprovider > /tmp/cv_kernels_tsan.log 2>&1 || true
echo "tsan race check: $(tail -1 /tmp/cv_kernels_tsan.log) | findings: $(grep -c 'WARNING: ThreadSanitizer' /tmp/cv_kernels_tsan.log)"
if [ "$1" = "full" ]; then  # the WHOLE -m gpu suite (reader, arena, faults, two-device gather) with the kernel source on the shim and asynchronous streams, ASan+UBSan
  LIB=$(python tests/simt_emu/build.py address,undefined)
  CV_TEST_MOCK_CUDA_LIB=$LIB MOCK_CUDA_ASYNC=1 MOCK_CUDA_JITTER_US=400 CV_SIMT_EMU_THREADS=4 CV_SIMT_EMU_SMS=16 LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    ASAN_OPTIONS=detect_leaks=0:halt_on_error=0:detect_stack_use_after_return=0 UBSAN_OPTIONS=print_stacktrace=1 \
    timeout 3000 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not read_to_tensor" -n 4 > /tmp/cv_gpu_suite_asan.log 2>&1 || true
  echo "asan+ubsan, whole GPU suite on the shim: $(tail -1 /tmp/cv_gpu_suite_asan.log) | findings: $(grep -ciE 'runtime error|ERROR: AddressSanitizer' /tmp/cv_gpu_suite_asan.log)"
fi
