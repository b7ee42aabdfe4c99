# ncu launch list of steady-state Mean-Teacher steps (eager launches; durations + DRAM bytes per launch), summarised per kernel and
# per family.  usage (GPU box): bash profiles/ncu_step.sh [tag]      -> gpurun_out/<tag>_launches.csv / _summary.txt
tag=${1:-r02}
python bench.py --steps 2 --warmup 3 --eager --no-cpu --quick > gpurun_out/${tag}_ncu_plain.json 2> gpurun_out/${tag}_ncu_plain.err || exit 1
L=$(python -c "
import json
d=json.loads([l for l in open('gpurun_out/${tag}_ncu_plain.json') if l.startswith('{')][-1])
print(d['gpu_launches']//d['steps'])")
echo "launches per step: $L"
# skip the 3 warm-up steps, capture the next 2
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $((3*L)) -c $((2*L)) --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu --quick > gpurun_out/${tag}_ncu_run.log 2>&1
python profiles/launch_summary.py gpurun_out/${tag}_launches.csv 2 > gpurun_out/${tag}_launch_summary.txt
head -70 gpurun_out/${tag}_launch_summary.txt; tail -12 gpurun_out/${tag}_launch_summary.txt
