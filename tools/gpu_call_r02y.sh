set -x
mkdir -p gpurun_out
WORKLOAD=config5 ITERS=4 timeout 900 ncu --set full --import-source on --clock-control none -k regex:wave -s 2 -c 2 -o gpurun_out/r02y_wave_c5 -f python tools/ab_frame.py wavefront=2 > gpurun_out/r02y_ncu.log 2>&1
tail -3 gpurun_out/r02y_ncu.log
