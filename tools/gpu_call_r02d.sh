set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_materials.py tests/test_gpu_errors_edges.py tests/test_gpu_median.py tests/test_gpu_models.py -q -m gpu > gpurun_out/r02d_pytest.log 2>&1
tail -15 gpurun_out/r02d_pytest.log
# dyn kernel vs single-ray kernel: configs 1, 2, 4 (+ config 3 at 1 spp and dyn=2 against the sample packets at 4 spp)
AB_DEVICE=1 AB_CLONES=1 AB_W=640 AB_H=480 AB_SPP=1 timeout 300 python tools/ab_render.py dyn=0,1 > gpurun_out/r02d_ab_dyn_c1.txt 2>&1
AB_DEVICE=1 AB_CLONES=1 AB_W=1920 AB_H=1080 AB_SPP=1 timeout 300 python tools/ab_render.py dyn=0,1 > gpurun_out/r02d_ab_dyn_c2.txt 2>&1
AB_DEVICE=1 AB_SPP=1 timeout 300 python tools/ab_render.py dyn=0,1 > gpurun_out/r02d_ab_dyn_c3spp1.txt 2>&1
AB_DEVICE=1 timeout 300 python tools/ab_render.py dyn=1,2 > gpurun_out/r02d_ab_dyn_c3.txt 2>&1
cat gpurun_out/r02d_ab_dyn_*.txt
WORKLOAD=config4 ITERS=8 timeout 600 python tools/ab_frame.py dyn=0,1 > gpurun_out/r02d_ab_dyn_c4.txt 2>&1
cat gpurun_out/r02d_ab_dyn_c4.txt
WORKLOAD=config5 ITERS=5 timeout 600 python tools/ab_frame.py packet=0,1 > gpurun_out/r02d_ab_packet_c5.txt 2>&1
cat gpurun_out/r02d_ab_packet_c5.txt
# why is a rank's kernel at world 8 twice the ideal? cold L2 (flush on/off), tile height, world
for fl in 1 0; do FLUSH=$fl PER_RANK=1 timeout 600 python tools/rank_sim.py 8 8,64,272 >> gpurun_out/r02d_rank_sim.txt 2>&1; done
FLUSH=1 timeout 600 python tools/rank_sim.py 2 8 >> gpurun_out/r02d_rank_sim.txt 2>&1
FLUSH=1 timeout 600 python tools/rank_sim.py 4 8 >> gpurun_out/r02d_rank_sim.txt 2>&1
cat gpurun_out/r02d_rank_sim.txt
