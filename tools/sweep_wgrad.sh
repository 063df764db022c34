# sweep of the weight-gradient kernel's per-stage cost weights (narrow = gamma(x) jobs, views = views-layer job)
# usage: bash tools/sweep_wgrad.sh RAYS "narrow values" "views values"
R=${1:-4096}
for n in ${2:-41 46 50 56}; do for v in ${3:-65 73 81}; do
  echo -n "rays=$R narrow=$n views=$v: "; NERFB_WG_NARROW_W=$n NERFB_WG_VIEWS_W=$v python tools/time_train.py $R 20 --no-fused | grep "^wgrad" | awk '{print $2, $3}'
done; done
