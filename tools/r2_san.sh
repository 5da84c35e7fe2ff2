# compute-sanitizer runs of the config-1 smoke (operator variants 0/1/7/8 + PCG incl. the triangular sweeps)
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|error" gpurun_out/r02_sanitizer_$tool.log | head -8
done
python -c "import __graft_entry__ as g; g.smoke()"
