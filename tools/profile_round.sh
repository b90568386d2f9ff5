set -x
python bench.py --steps 50 --warmup 3 > gpurun_out/s5_bench_full.json 2> gpurun_out/s5_bench_full.err
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/s5_bench_ref.json 2> gpurun_out/s5_bench_ref.err
python tools/run_configs.py --out gpurun_out/s5_configs.json > gpurun_out/s5_configs.log 2>&1
python tools/primitives_bench.py > gpurun_out/s5_primitives.json 2> gpurun_out/s5_primitives.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_s5.csv python bench.py --steps 15 --warmup 3 --no-cpu > gpurun_out/ncu_s5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'tile_kernel|sa_fwd_kernel|sa_bwd_kernel|segsum_kernel|gemm_tc_kernel' --launch-skip 120 --launch-count 30 -f -o gpurun_out/prof_s5_fused python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_s5_full.log 2>&1
ls -la gpurun_out/prof_s5_fused.ncu-rep
