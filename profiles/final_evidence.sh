# Round-2 evidence run on one GPU: full GPU test suite, bench line of every BASELINE configuration, reference arm, smoke.
O=gpurun_out/final
mkdir -p $O
rm -f gpurun_out/full_size_report.txt
python -m pytest tests -m gpu -q > $O/gpu_tests.txt 2>&1; tail -3 $O/gpu_tests.txt
cp gpurun_out/full_size_report.txt $O/full_size_parity_report.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
python bench.py --steps 20 --warmup 5 > $O/bench_mt_acdc.json 2> $O/bench_mt_acdc.err
[ -n "$SKIP_REF" ] || python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python profiles/hpfg_step_throughput.py > $O/hpfg_step.txt 2>&1; tail -2 $O/hpfg_step.txt
for c in mt_cfg1 mt_isic cps uamt; do python bench.py --config $c --steps 20 --warmup 5 --quick > $O/bench_$c.json 2> $O/bench_$c.err; done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "no json", e); continue
    print("%-40s %9.1f img/s %8.3f ms  e2e %9.1f  cpu %s" % (f.split("/")[-1], d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("cpu_baseline") or {}).get("value")))
PY
