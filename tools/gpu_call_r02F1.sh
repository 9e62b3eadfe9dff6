# round 2, final single-GPU evidence: tests, sanitizers, the three bench workloads, the reference arm, ncu launch list + full captures
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r02F_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02F_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02F_bench_config3_1gpu.json 2> gpurun_out/r02F_bench_config3_1gpu.err; tail -2 gpurun_out/r02F_bench_config3_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02F_reference_arm.json 2> gpurun_out/r02F_reference_arm.err
timeout 600 python bench.py --workload config4 --steps 5 --warmup 6 > gpurun_out/r02F_bench_config4_1gpu.json 2> gpurun_out/r02F_bench_config4_1gpu.err; tail -2 gpurun_out/r02F_bench_config4_1gpu.err
timeout 600 python bench.py --workload config5 --steps 5 --warmup 3 > gpurun_out/r02F_bench_config5_1gpu.json 2> gpurun_out/r02F_bench_config5_1gpu.err; tail -2 gpurun_out/r02F_bench_config5_1gpu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02F_launches_bench_1gpu.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02F_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_packet -s 12 -c 1 -f -o gpurun_out/r02F_render_packet python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02F_ncu_packet.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 12 -c 1 -f -o gpurun_out/r02F_config4_render python bench.py --workload config4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02F_ncu_config4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wave -s 12 -c 2 -f -o gpurun_out/r02F_config5_wave python bench.py --workload config5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02F_ncu_config5.log 2>&1
ls -la gpurun_out/r02F_*
timeout 300 python tools/run_configs.py 1,2 gpurun_out/r02F_configs12.json | cut -c1-300
timeout 120 python tools/sanitize_smoke.py 2>&1 | tail -3
SAN_TIMEOUT=400 bash tools/sanitize.sh 2>&1 | tail -20
