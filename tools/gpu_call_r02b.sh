set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02b_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02b_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench_config3.json 2> gpurun_out/r02b_bench_config3.err
tail -c 3000 gpurun_out/r02b_bench_config3.json; tail -5 gpurun_out/r02b_bench_config3.err
timeout 600 python bench.py --workload config4 --steps 5 --warmup 3 > gpurun_out/r02b_bench_config4.json 2> gpurun_out/r02b_bench_config4.err
tail -c 2500 gpurun_out/r02b_bench_config4.json; tail -5 gpurun_out/r02b_bench_config4.err
timeout 600 python bench.py --workload config5 --steps 3 --warmup 3 > gpurun_out/r02b_bench_config5.json 2> gpurun_out/r02b_bench_config5.err
tail -c 2500 gpurun_out/r02b_bench_config5.json; tail -5 gpurun_out/r02b_bench_config5.err
