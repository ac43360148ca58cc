#!/bin/bash
mkdir -p gpurun_out
{
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize.py 2>&1 | grep -v "^=========     " | tail -14
done
} > gpurun_out/compute_sanitizer.log 2>&1
tail -34 gpurun_out/compute_sanitizer.log
