timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1
grep -v "^=========     at\|^=========         in\|Host Frame\|^=========$" gpurun_out/sanitizer.log | head -40
