# A/B of an environment switch on one box, back to back: profiles/ab_env.sh VAR value1 value2 ...
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 20 --warmup 5 --no-cpu --quick 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
k=d['kernel_time_per_step']
print('$var=$v : %.3f ms/step  %.0f img/s  (e2e %.3f ms)  serialized: conv %.3f glue %.3f wgrad %.3f' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], k['conv_tcgen05']['ms_per_step'], k['bn_pool_upsample_glue']['ms_per_step'], k['wgrad_tcgen05']['ms_per_step']))"
done
