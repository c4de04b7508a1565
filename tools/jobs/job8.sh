nproc; free -g | head -2
B="--no-cpu-baseline --no-e2e --no-fastq --no-sharded"
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 $B ${@:3} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', d['ms_per_step'], d['value'], d['clocks'].get('source'), d['clocks'].get('samples'))"; }
run 29511 nvml
BK_BENCH_SAMPLER=smi run 29512 smi
run 29513 nvml_inflight3 --in-flight 3
run 29514 nvml_inflight6 --in-flight 6
