#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitizer_smoke5.py > gpurun_out/sanitizer5_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|balloon|solid|multipatch|Error|error" gpurun_out/sanitizer5_$tool.log | head -20
done
