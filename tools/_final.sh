timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/t_final_bf16.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_final_bf16.json 2> gpurun_out/bench_r02_final_bf16.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_bf16.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_bf16_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/ncu_list_bf16.log 2>&1
timeout 200 python tools/bench_emission_variants.py > gpurun_out/variants_bf16.json 2> gpurun_out/variants_bf16.err
timeout 250 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/cfg3_final_bf16.json 2> gpurun_out/cfg3_final_bf16.err
timeout 250 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/cfg5_final_bf16.json 2> gpurun_out/cfg5_final_bf16.err
