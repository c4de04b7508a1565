B="--no-cpu-baseline --no-e2e --no-fastq --no-sharded"
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02.csv python bench.py $B --steps 4 --warmup 3 --in-flight 1 > gpurun_out/bench_under_ncu_r02.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --launch-skip 164 -c 41 -o gpurun_out/r02_full -f python bench.py $B --steps 2 --warmup 3 --in-flight 1 > gpurun_out/ncu_r02_full.log 2>&1
tail -2 gpurun_out/ncu_r02_full.log | cut -c1-200; ls -la gpurun_out/r02_full.ncu-rep
