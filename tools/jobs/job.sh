B="--no-cpu-baseline --no-e2e --no-fastq --no-sharded"
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 || exit 1
timeout 400 python -m pytest tests -x -q -m gpu --timeout 60 -k "not 200 and not fullsize" 2>&1 | tail -8
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu --timeout 200 -k "not 200" 2>&1 | tail -4
timeout 120 python bench.py $B --in-flight 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('inflight 4', d['ms_per_step'], d['stage_ms_single_sample'], d['latency_ms_single_sample'])"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/launches_r2l_warm.csv python bench.py $B --steps 4 --warmup 3 --in-flight 1 > /dev/null 2>&1
python tools/launch_share.py gpurun_out/launches_r2l_warm.csv 2>&1 | tail -45
