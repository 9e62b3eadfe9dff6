# round 2, final: ncu --set full of the config 3 / config 4 render kernels inside the resident warm-up frames of bench.py (6th frame)
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_packet -s 5 -c 1 -f -o gpurun_out/r02F_render_packet python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02F_ncu_packet.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 5 -c 1 -f -o gpurun_out/r02F_config4_render python bench.py --workload config4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02F_ncu_config4.log 2>&1
ls -la gpurun_out/r02F_*.ncu-rep
