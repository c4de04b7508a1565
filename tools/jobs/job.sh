timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 || exit 1
S=$(date +%s); timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -4; echo "pytest wall $(( $(date +%s) - S ))s"
S=$(date +%s); timeout 600 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; echo "bench rc=$? wall $(( $(date +%s) - S ))s"
S=$(date +%s); timeout 300 python bench.py --impl reference > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err; echo "ref rc=$? wall $(( $(date +%s) - S ))s"
cut -c1-250 gpurun_out/bench_r02_n1.json; cut -c1-400 gpurun_out/bench_r02_ref.json
