B="--no-cpu-baseline --no-e2e --no-fastq --no-sharded"
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 || exit 1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -m gpu --timeout 60 2>&1 | tail -3
for f in 4 6 8; do timeout 120 python bench.py $B --in-flight $f 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('inflight $f', d['ms_per_step'], 'cpu_ms', d['host_cpu_ms_per_step']['user'], d['host_cpu_ms_per_step']['sys'], 'lat', d['latency_ms_single_sample'])"; done
BK_SPIN=1 timeout 120 python bench.py $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spin', d['ms_per_step'], 'cpu_ms', d['host_cpu_ms_per_step']['user'], d['host_cpu_ms_per_step']['sys'], 'lat', d['latency_ms_single_sample'])"
