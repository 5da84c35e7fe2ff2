# compute-sanitizer runs of the config-1 smoke (operator variants 0/1/7/9 + PCG incl. the DMMA triangular sweeps)
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" gpurun_out/r02_sanitizer_$tool.log | head -4
  grep -E "Race reported|Invalid|Error:" gpurun_out/r02_sanitizer_$tool.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -12
done
