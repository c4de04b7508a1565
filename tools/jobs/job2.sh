timeout 200 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu --timeout 120 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --no-fastq --no-cpu-baseline --shard-depth 400000 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
sh=d['sharded']; print('sharded', sh['ms_per_sample'], sh['value'], sh.get('bit_equal'), sh['collectives'], sh.get('unsharded_same_sample_ms_rank0'))
"
