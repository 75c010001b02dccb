#!/bin/bash
# the three variants of vtb_input_batch: parity against the reference's golden batches + timings (torch-free harness)
mkdir -p gpurun_out
for v in "" --variant2 --variant3; do
  timeout 200 python tools/input_selftest.py $v > gpurun_out/input_selftest${v#--}.log 2>&1; echo "selftest $v exit=$?"
  grep -E "us|PASS|FAIL|bit" gpurun_out/input_selftest${v#--}.log | tail -8 | cut -c1-220
done
