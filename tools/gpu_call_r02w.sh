set -x
mkdir -p gpurun_out
AB_DEVICE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_packet -s 3 -c 1 -f -o gpurun_out/r02w_shadow_packet python tools/ab_render.py shadow_packets=1 shadows > gpurun_out/r02w_ncu.log 2>&1
tail -3 gpurun_out/r02w_ncu.log
