# A/B of the in-kernel compositing on one box: bench main leg with the path on / off
run() { # label, env...
  label=$1; shift
  env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train --no-c0 --no-llff > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$label.json').read().strip().splitlines()[-1]); print('$label', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['achieved'],1), round(d['roofline']['kernel_share_of_step'],4), d['clocks']['sm_mhz'])"
}
run off1 NERFB200_COMPOSITE_IN_KERNEL=0
run on1 X=1
run off2 NERFB200_COMPOSITE_IN_KERNEL=0
run on2 X=1
