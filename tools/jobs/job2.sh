timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 || exit 1
timeout 200 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_fastq.py -x -q -m gpu --timeout 120 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --no-fastq --no-cpu-baseline --shard-depth 200000 > gpurun_out/bench_r02_n2b.json 2> gpurun_out/bench_r02_n2b.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_n2b.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['latency_ms_single_sample'], d['clocks'])
sh=d['sharded']; print('sharded', sh['ms_per_sample'], sh.get('bit_equal'))
"
tail -3 gpurun_out/bench_r02_n2b.err
