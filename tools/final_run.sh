#!/bin/bash
# tools/final_run.sh TAG — the round's evidence in one GPU call: GPU tests, smoke, the bench line, the ncu launch list of
# the bench command and one `--set full` capture of the kernels of one sample (profiles/README.md).
cd "$(dirname "$0")/.."
tag=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cut -c1-700 gpurun_out/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --in-flight 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_scan|k_leftover|k_bin_hist|k_bin_scatter|k_bin_count|k_map_grp|k_noise_seq" \
    --launch-skip 26 --launch-count 13 -f -o gpurun_out/full_$tag python bench.py --steps 1 --warmup 3 --in-flight 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/full_$tag.ncu-rep gpurun_out/launches_$tag.csv
