#!/bin/bash
# e2e step time of bench.py for different RTDS_BANDS settings (row bands whose D2H overlaps later bands' rendering)
for b in ${BANDS_LIST:-1 2 3 4}; do
  RTDS_BANDS=$b python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bands', $b, d['e2e']['ms_per_step'], d['e2e']['value'])"
done
