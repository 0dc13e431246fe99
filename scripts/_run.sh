ncu --set full --clock-control none --import-source on -k regex:egn_mlp_tc -s 3 -c 1 -o gpurun_out/prof_mlp_tc_bf16 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rays 16384 --mlp tc_bf16 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:egn_gather_kernel -s 3 -c 1 -o gpurun_out/prof_gather python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rays 16384 --mlp tc_bf16 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:egn_coarse_kernel -s 3 -c 1 -o gpurun_out/prof_coarse python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rays 16384 --mlp tc_bf16 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu1.log; ls -la gpurun_out/*.ncu-rep
