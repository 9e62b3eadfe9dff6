set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02u_bench_config3_1gpu.json 2> gpurun_out/r02u_bench_config3_1gpu.err; tail -2 gpurun_out/r02u_bench_config3_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02u_reference_arm.json 2> gpurun_out/r02u_reference_arm.err
timeout 600 python bench.py --workload config4 --steps 5 --warmup 6 > gpurun_out/r02u_bench_config4_1gpu.json 2> gpurun_out/r02u_bench_config4_1gpu.err; tail -2 gpurun_out/r02u_bench_config4_1gpu.err
timeout 600 python bench.py --workload config5 --steps 3 --warmup 3 > gpurun_out/r02u_bench_config5_1gpu.json 2> gpurun_out/r02u_bench_config5_1gpu.err; tail -2 gpurun_out/r02u_bench_config5_1gpu.err
# launch list of the bench command (cold-cache, serialised per-launch times: the kernels' SHARES of a step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02u_launches_bench_1gpu.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02u_ncu_bench.log 2>&1
# ncu --set full of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_packet -s 6 -c 1 -f -o gpurun_out/r02u_render_packet python bench.py --steps 2 --warmup 6 --no-cpu-baseline > gpurun_out/r02u_ncu_packet.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mt_expand_dirs -s 6 -c 1 -f -o gpurun_out/r02u_dirs python bench.py --steps 2 --warmup 6 --no-cpu-baseline > gpurun_out/r02u_ncu_dirs.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 8 -c 1 -f -o gpurun_out/r02u_config4_render python bench.py --workload config4 --steps 2 --warmup 8 --no-cpu-baseline > gpurun_out/r02u_ncu_config4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_full -s 3 -c 1 -f -o gpurun_out/r02u_config5_render python bench.py --workload config5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02u_ncu_config5.log 2>&1
ls -la gpurun_out/r02u_*
timeout 300 python tools/run_configs.py 1,2 gpurun_out/r02u_configs12.json | cut -c1-300
