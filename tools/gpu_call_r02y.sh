set -x
mkdir -p gpurun_out
cd raytracer-data-structures_b200/csrc
for v in "128 6" "64 12" "64 10" "32 24"; do
  set -- $v
  touch render.cu
  make -j16 EXTRA="-DRTDS_PK_THREADS=$1 -DRTDS_PK_MINB=$2" 2>&1 | grep -E "error" -A3
  (cd ../..; echo "== PK_THREADS=$1 MINB=$2"; ITERS=12 timeout 600 python tools/ab_frame.py lpt=1 2>&1 | cut -c1-200 | head -1; WORLD=8 ITERS=10 timeout 600 python tools/ab_frame.py lpt=1 2>&1 | cut -c1-200 | head -1)
done
