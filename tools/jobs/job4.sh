S=$(date +%s)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 > gpurun_out/bench_r02_n4.json 2> gpurun_out/bench_r02_n4.err
echo rc=$? wall=$(( $(date +%s) - S ))s
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_n4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['latency_ms_single_sample'], d['e2e']['value'], d['gpu_launches'], d['clocks']['samples'])
sh=d['sharded']; print('sharded', sh['ms_per_sample'], sh['value'], sh.get('bit_equal'), sh['collectives']['ms_per_sample_rank0'])
print('fastq', d['fastq']['bgzf']['samples_per_min'], d['fastq']['plain_gz']['samples_per_min'])
"
tail -2 gpurun_out/bench_r02_n4.err
