B="--no-cpu-baseline --no-e2e --no-fastq --no-sharded"
for f in 4 6 4 6; do timeout 60 python bench.py $B --in-flight $f 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('inflight $f', round(d['ms_per_step'],4), d['host_cpu_ms_per_step']['user'], d['host_cpu_ms_per_step']['sys'])"; done
