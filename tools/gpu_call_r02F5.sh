# round 2 final pass on the last build: GPU suite, sanitizers, the three bench lines, the reference arm, smoke
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r02F_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02F_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02F_bench_config3_1gpu.json 2> gpurun_out/r02F_bench_config3_1gpu.err; tail -2 gpurun_out/r02F_bench_config3_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02F_reference_arm.json 2> gpurun_out/r02F_reference_arm.err
timeout 600 python bench.py --workload config4 --steps 5 --warmup 6 > gpurun_out/r02F_bench_config4_1gpu.json 2> gpurun_out/r02F_bench_config4_1gpu.err; tail -2 gpurun_out/r02F_bench_config4_1gpu.err
timeout 600 python bench.py --workload config5 --steps 5 --warmup 3 > gpurun_out/r02F_bench_config5_1gpu.json 2> gpurun_out/r02F_bench_config5_1gpu.err; tail -2 gpurun_out/r02F_bench_config5_1gpu.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
timeout 120 python tools/sanitize_smoke.py 2>&1 | tail -2
SAN_TIMEOUT=400 bash tools/sanitize.sh 2>&1 | grep -E "exit|SUMMARY"
