#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/${1:-r2}_sanitizers.txt
: > $out
for tool in memcheck racecheck initcheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_case.py" >> $out
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_case.py 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize_case|SUMMARY|Invalid|hazard|Error|error" | head -20 >> $out
done
cat $out
