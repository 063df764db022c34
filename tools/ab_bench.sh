# repeated A/B of experiment builds on the bench main leg (one box): bash tools/ab_bench.sh "" h11 "" h11 ...
for v in "$@"; do
  NERFB_VARIANT=$v python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-train --no-c0 --no-llff 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench [$v]', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'])"
done
