run() { label=$1; shift
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train --no-c0 --no-llff "$@" > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$label.json').read().strip().splitlines()[-1]); print('$label', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['achieved'],1), round(d['roofline']['kernel_share_of_step'],4), d['clocks']['sm_mhz'])"
}
run c32k_1 --chunk 32768
run c640k_1 --chunk 640000
run c32k_2 --chunk 32768
run c160k_2 --chunk 160000
