python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
ncu --set full --clock-control none --import-source on -k regex:egn_fused -s 3 -c 1 -o gpurun_out/prof_fused_bf16_65536 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-line > gpurun_out/ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_fused.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity-line > gpurun_out/ncu2.log 2>&1
