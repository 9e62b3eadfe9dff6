# round 2 final: bench lines again after the side measurements got their own warm-up (with_shadows, e2e variants)
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02F_bench_config3_1gpu.json 2> gpurun_out/r02F_bench_config3_1gpu.err; tail -2 gpurun_out/r02F_bench_config3_1gpu.err
timeout 600 python bench.py --workload config4 --steps 5 --warmup 6 > gpurun_out/r02F_bench_config4_1gpu.json 2> gpurun_out/r02F_bench_config4_1gpu.err; tail -2 gpurun_out/r02F_bench_config4_1gpu.err
timeout 600 python bench.py --workload config5 --steps 5 --warmup 3 > gpurun_out/r02F_bench_config5_1gpu.json 2> gpurun_out/r02F_bench_config5_1gpu.err; tail -2 gpurun_out/r02F_bench_config5_1gpu.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
