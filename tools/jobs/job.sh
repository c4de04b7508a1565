S=$(date +%s); timeout 300 python bench.py > gpurun_out/bench_r02_n1_final.json 2> gpurun_out/bench_r02_n1_final.err; echo "bench rc=$? wall $(( $(date +%s) - S ))s"
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_n1_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['latency_ms_single_sample'], d['clocks']['samples'], d['e2e']['value'], d['fastq']['bgzf']['samples_per_min'], d['sharded']['ms_per_sample'])
"
B="--no-cpu-baseline --no-e2e --no-fastq --no-sharded"
for i in 1 2; do timeout 100 python bench.py $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('short run', d['ms_per_step'], d['latency_ms_single_sample'])"; done
