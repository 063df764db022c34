# A/B of experiment builds (NERFB_VARIANT=name -> libnerf_b200_<name>.so) on one box: bench main leg + the training kernels
# usage: bash tools/ab_variants.sh "" tw2 tw20 ""
for v in "$@"; do
  NERFB_VARIANT=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train --no-c0 --no-llff 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench [$v]', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'])"
  NERFB_VARIANT=$v python tools/time_train.py 4096 10 2>&1 | grep -v fused | awk -v v="$v" '{print "   [" v "] " $0}'
done
