S=$(date +%s)
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/bench_r02_n8.json 2> gpurun_out/bench_r02_n8.err
echo rc=$? wall=$(( $(date +%s) - S ))s
cut -c1-200 gpurun_out/bench_r02_n8.json; tail -3 gpurun_out/bench_r02_n8.err
