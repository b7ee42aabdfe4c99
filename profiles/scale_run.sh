# Data-parallel bench lines on N GPUs of one box: bash profiles/scale_run.sh N config...   -> gpurun_out/final/bench_<config>_<N>gpu.json
N=$1; shift
mkdir -p gpurun_out/final
port=29600
for c in "$@"; do
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --config $c \
      > gpurun_out/final/bench_${c}_${N}gpu.json 2> gpurun_out/final/bench_${c}_${N}gpu.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/final/bench_${c}_${N}gpu.json") if l.startswith("{")][-1])
    print("$c N=$N: %.0f img/s  %.3f ms/step  e2e %.0f  dp_check %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: v for k, v in d["dp_check"].items() if k != "note"}))
except Exception as e:
    print("$c N=$N: no line (%s)" % e)
PY
done
