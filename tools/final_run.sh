#!/bin/bash
# tools/final_run.sh TAG — the round's evidence in one GPU call (run it through gpurun on ONE B200): smoke, GPU tests, the
# bench line and the reference arm, the ncu launch list of the bench command and one `--set full` capture of the 41
# launches of one sample (profiles/README.md).  Every step has its own time limit: a hang costs minutes, not the box.
cd "$(dirname "$0")/.."
tag=${1:-final}
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --no-fastq --no-sharded"
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 || exit 1
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cut -c1-400 gpurun_out/bench_$tag.json
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; cut -c1-300 gpurun_out/bench_ref_$tag.json
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py $B --steps 4 --warmup 3 --in-flight 1 > gpurun_out/bench_under_ncu_$tag.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on --launch-skip 164 -c 41 -f -o gpurun_out/full_$tag \
    python bench.py $B --steps 2 --warmup 3 --in-flight 1 > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/full_$tag.ncu-rep gpurun_out/launches_$tag.csv
# back in the container:  python profiles/summarize.py gpurun_out/full_$tag.ncu-rep gpurun_out/launches_$tag.csv rNN
#                         python tools/src_hot.py gpurun_out/full_$tag.ncu-rep k_scan 8
