timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 || exit 1
S=$(date +%s); timeout 200 python bench.py > gpurun_out/bench_r02_n1_if6.json 2> gpurun_out/bench_r02_n1_if6.err; echo "bench rc=$? wall $(( $(date +%s) - S ))s"
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_n1_if6.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['latency_ms_single_sample'], d['clocks']['samples'], d['e2e']['value'], d['fastq']['bgzf']['samples_per_min'], d['fastq']['plain_gz']['samples_per_min'], d['sharded']['ms_per_sample'], d['config']['samples_in_flight_per_gpu'])
"
tail -2 gpurun_out/bench_r02_n1_if6.err
