set -x
mkdir -p gpurun_out
AB_DEVICE=1 timeout 900 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:'wave|render_packet' --csv --log-file gpurun_out/r02y_wave_c3.csv python tools/ab_render.py wavefront=0,1 shadows > gpurun_out/r02y_c3.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02y_wave_c3.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[iid], r[ik][:40]),{})[r[im]]=r[iv]
for k,v in list(d.items())[-9:]: print(k, v)
PY
