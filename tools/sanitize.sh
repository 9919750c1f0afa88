#!/bin/bash
# compute-sanitizer memcheck + racecheck over smoke() and one parity test per kernel family (run on the GPU box:
#   gpurun --timeout 2400 -- 'bash tools/sanitize.sh').  Logs go to gpurun_out/ (copied to profiles/ by hand).
set -u
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
# the sync kernels (warp / TMA / plain-load), the three frame kernels (warp per frame at 64 / 1024 / 2048, CTA per
# frame, any-fft_len), TX (warp per packet / generic / rolloff), chain kernels, conditioning (AGC, IIR, PAPR), CRC
TESTS="tests/test_gpu_parity.py::test_rx_parity_c1 tests/test_gpu_parity.py::test_rx_parity_c3 \
tests/test_gpu_parity.py::test_warp_frame_kernel_fft2048 tests/test_gpu_parity.py::test_sync_kernel_variants_detect_bits \
tests/test_gpu_parity.py::test_frame_kernel_variants_agree tests/test_gpu_parity.py::test_tx_warp_kernel_ragged \
tests/test_gpu_parity.py::test_rx_truncated_and_corrupt tests/test_gpu_parity.py::test_crc32_parity \
tests/test_round2.py::test_several_carrier_sets_and_pilot_inside_occupied tests/test_round2.py::test_oversize_frame_is_consumed_and_the_stream_goes_on \
tests/test_round2.py::test_unaligned_payload_slots_take_the_byte_path \
tests/test_next_rows.py::test_agc2_parity tests/test_next_rows.py::test_iir_ccd_parity tests/test_next_rows.py::test_papr_sink_gpu \
tests/test_next_rows.py::test_tx_rolloff_parity_and_loopback tests/test_next_rows.py::test_runtime_reconfiguration"
for tool in memcheck racecheck; do
  log=$OUT/sanitizer_${tool}.log
  echo "== $tool: smoke()" > $log
  timeout 900 $CS --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" >> $log 2>&1
  echo "rc=$?" >> $log
  echo "== $tool: parity tests" >> $log
  timeout 1500 $CS --tool $tool --print-limit 20 python -m pytest -x -q -m gpu $TESTS >> $log 2>&1
  echo "rc=$?" >> $log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" $log | tail -8
done
