for v in "6 100" "6 85" "6 70" "8 85" "8 70" "6 100"; do set -- $v
  RTDS_BANDS=$1 RTDS_BAND_RATIO=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bands', $1, 'ratio', $2, d['e2e']['ms_per_step'], d['e2e']['value'])"
done
