#!/bin/bash
# compute-sanitizer over the hot path (run under gpurun): memcheck and racecheck on the smoke batch and on a reduced
# set of the parity tests (every decoder: speculative tokenizer NW = 1 / 4, LZ executor bytes / symbols / wide,
# fallback decoder, CRC fold, STORE copy, method-93 container, Zstandard tokenizer, compressor), logs under profiles/.
#   usage: tools/sanitize.sh <tag>
T=${1:-r2}
mkdir -p gpurun_out
export OTZ_SANITIZE=1
TESTS="tests/test_gpu_golden.py tests/test_gpu_crc.py tests/test_gpu_extract.py tests/test_gpu_inflate_twophase.py::test_shapes_match_oracle tests/test_gpu_inflate_twophase.py::test_stored_blocks_stay_on_the_fast_path tests/test_gpu_inflate_twophase.py::test_packed_output_arena_any_alignment tests/test_gpu_zstd.py tests/test_gpu_deflate.py tests/test_gpu_round2.py::test_pipelined_host_call_equals_one_batch tests/test_gpu_round2.py::test_multi_device_call_equals_single_device_call"
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/${T}_sanitizer_${tool}_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_sanitizer_${tool}_smoke.out 2>&1
  echo "$tool smoke rc=$?" >> gpurun_out/${T}_sanitizer_summary.txt
  timeout ${SAN_TIMEOUT:-1500} compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/${T}_sanitizer_${tool}_tests.log python -m pytest $TESTS -x -q > gpurun_out/${T}_sanitizer_${tool}_tests.out 2>&1
  echo "$tool tests rc=$?" >> gpurun_out/${T}_sanitizer_summary.txt
  tail -3 gpurun_out/${T}_sanitizer_${tool}_tests.out >> gpurun_out/${T}_sanitizer_summary.txt
  grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/${T}_sanitizer_${tool}_*.log >> gpurun_out/${T}_sanitizer_summary.txt
done
cat gpurun_out/${T}_sanitizer_summary.txt
