#!/bin/bash
# The GPU ingest pipeline's host logic under sanitizers, without a GPU: tests/test_gpu_reader.py against the mock CUDA
# runtime of tests/mock_cuda (see tests/test_ingest_pipeline_cpu.py), built with -fsanitize=thread, then address,undefined.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"; cd "$ROOT"
K="not read_to_tensor"
LIB=$(python tests/mock_cuda/build.py thread)
CV_TEST_MOCK_CUDA_LIB=$LIB LD_PRELOAD="$(gcc -print-file-name=libtsan.so)" TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4" \
  setarch "$(uname -m)" -R python -m pytest tests/test_gpu_reader.py tests/test_zz_gpu_reader_faults.py tests/test_arena_gpu.py tests/test_gds_gpu.py -m gpu -q -p no:cacheprovider -k "$K" > /tmp/cv_ingest_tsan.log 2>&1 || true
echo "tsan: $(tail -1 /tmp/cv_ingest_tsan.log) | findings: $(grep -c 'WARNING: ThreadSanitizer' /tmp/cv_ingest_tsan.log)"
LIB=$(python tests/mock_cuda/build.py address,undefined)
CV_TEST_MOCK_CUDA_LIB=$LIB LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 \
  UBSAN_OPTIONS=print_stacktrace=1 python -m pytest tests/test_gpu_reader.py tests/test_zz_gpu_reader_faults.py tests/test_arena_gpu.py tests/test_gds_gpu.py -m gpu -q -p no:cacheprovider -k "$K" > /tmp/cv_ingest_asan.log 2>&1 || true
echo "asan+ubsan: $(tail -1 /tmp/cv_ingest_asan.log) | findings: $(grep -ciE 'runtime error|AddressSanitizer' /tmp/cv_ingest_asan.log)"
# Stream-order check: the whole -m gpu suite with the kernel source on the SIMT shim (tests/simt_emu) and the runtime stand-in's streams running
# ASYNCHRONOUSLY (a thread per stream, random pauses up to 1 ms, ordered only by events).  A forgotten dependency = wrong bytes in a parity test.
# First the planted one (the verify stream no longer waits for a group's copies): it must fail; then the product as it is: it must pass.
for M in verifier_does_not_wait_for_the_copy copies_do_not_wait_for_the_callers_stream callers_stream_does_not_wait_for_the_read; do
  LIB=$(python tests/simt_emu/build.py --mutate $M)
  CV_TEST_MOCK_CUDA_LIB=$LIB MOCK_CUDA_ASYNC=1 MOCK_CUDA_JITTER_US=500 CV_SIMT_EMU_THREADS=3 timeout 1200 python -m pytest tests/test_gpu_reader.py tests/test_arena_gpu.py \
    tests/test_zzz_stream_order_gpu.py -m gpu -q -p no:cacheprovider -k "$K" -n 6 > /tmp/cv_ingest_async_planted.log 2>&1 || true
  echo "async streams, planted: $M: $(tail -1 /tmp/cv_ingest_async_planted.log) (want failures)"
done
LIB=$(python tests/simt_emu/build.py)
for J in 300 1000 2000; do
  CV_TEST_MOCK_CUDA_LIB=$LIB MOCK_CUDA_ASYNC=1 MOCK_CUDA_JITTER_US=$J CV_SIMT_EMU_THREADS=3 timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -k "$K" -n 6 > /tmp/cv_ingest_async.log 2>&1 || true
  echo "async streams, jitter $J us: $(tail -1 /tmp/cv_ingest_async.log)"
done
# TSan on top of that: the same library built with -fsanitize=thread (kernel source as TSan fibers, tools/sanitize_kernels.sh), streams asynchronous.
# A pinned-ring slot refilled by a fetch thread while its H2D copy is still reading it, a kernel consuming a buffer another stream is writing: data
# races between host threads, stream threads and CUDA threads, reported with both stacks.  Planted bug first (must be reported), then the product.
ARCH="$(uname -m)"
TS="halt_on_error=0 report_signal_unsafe=0 history_size=4"
LIB=$(python tests/simt_emu/build.py --mutate verifier_does_not_wait_for_the_copy thread)
CV_TEST_MOCK_CUDA_LIB=$LIB MOCK_CUDA_ASYNC=1 MOCK_CUDA_JITTER_US=200 CV_SIMT_EMU_THREADS=2 CV_SIMT_EMU_SMS=6 LD_PRELOAD="$(gcc -print-file-name=libtsan.so)" TSAN_OPTIONS="$TS" \
  timeout 1500 setarch $ARCH -R python -m pytest tests/test_gpu_reader.py -m gpu -q -p no:cacheprovider -k "c1_file" -n 3 > /tmp/cv_ingest_tsan_async_planted.log 2>&1 || true
echo "tsan + async streams, planted: $(tail -1 /tmp/cv_ingest_tsan_async_planted.log) | races reported: $(grep -c 'ThreadSanitizer: data race' /tmp/cv_ingest_tsan_async_planted.log) (want > 0)"
LIB=$(python tests/simt_emu/build.py thread)
CV_TEST_MOCK_CUDA_LIB=$LIB MOCK_CUDA_ASYNC=1 MOCK_CUDA_JITTER_US=200 CV_SIMT_EMU_THREADS=2 CV_SIMT_EMU_SMS=6 LD_PRELOAD="$(gcc -print-file-name=libtsan.so)" TSAN_OPTIONS="$TS" \
  timeout 3300 setarch $ARCH -R python -m pytest tests/test_gpu_reader.py tests/test_arena_gpu.py tests/test_zz_gpu_reader_faults.py tests/test_zzz_stream_order_gpu.py tests/test_zzz_hostile_gpu.py \
  -m gpu -q -p no:cacheprovider -k "$K" -n 3 > /tmp/cv_ingest_tsan_async.log 2>&1 || true
echo "tsan + async streams: $(tail -1 /tmp/cv_ingest_tsan_async.log) | findings: $(grep -c 'WARNING: ThreadSanitizer' /tmp/cv_ingest_tsan_async.log)"
