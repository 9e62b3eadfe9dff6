set -x
mkdir -p gpurun_out
cat > /tmp/sah7m.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tools'))
import run_configs as RC
rt = RC.rt
ctx = rt.Rtds(0)
sph, mat = RC.torus_knot_scene(7_000_000)
ctx.set_spheres(sph, mat)
for i in range(2):
    st = ctx.build(rt.BVH, mode=rt.MODE_SAH)
    print("SAH 7M: %.2f ms, %d launches, depth %d" % (st["ms"], st["kernel_launches"], st["max_depth"]))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02o_sah7m_launches.csv python /tmp/sah7m.py > gpurun_out/r02o_sah7m.log 2>&1
tail -3 gpurun_out/r02o_sah7m.log
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r02o_sah7m_launches.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    us = v / 1000 if u in ('nsecond', 'ns') else (v * 1000 if u in ('msecond', 'ms') else v)
    name = r[ki].split('(')[0].split('::')[-1]
    tot[name][0] += 1; tot[name][1] += us
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-40s %5d launches %10.1f us" % (k, c, t))
PY
