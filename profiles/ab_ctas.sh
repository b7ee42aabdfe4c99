# A/B of the persistent-CTA budgets per kernel kind (forward convs / data gradients / weight gradients), one box, back to back
for cfg in "148 148 148" "74 148 148" "74 74 74" "74 100 48" "148 100 48" "100 100 48" "74 110 36" "148 120 28"; do
  set -- $cfg
  HPFG_CTAS_FWD=$1 HPFG_CTAS_DGRAD=$2 HPFG_CTAS_WGRAD=$3 python bench.py --steps 20 --warmup 5 --no-cpu --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ctas fwd/dgrad/wgrad $cfg : %.3f ms/step  %.0f img/s  (e2e %.3f ms)' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step']))"
done
