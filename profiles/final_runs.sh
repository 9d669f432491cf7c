# End-of-round check on one GPU box: GPU tests, smoke, bench lines of the headline and of BASELINE configs 1-5, reference arm,
# ncu launch list of the default bench command, workload family timings.   bash profiles/final_runs.sh   (outputs in gpurun_out/)
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_gpu_tests.log
python bench.py > gpurun_out/final_bench_headline.json 2> gpurun_out/final_bench_headline.err
for c in 1 2 3 4 5; do python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/final_bench_config_$c.json 2> gpurun_out/final_bench_config_$c.err; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_reference_arm.json 2> gpurun_out/final_reference_arm.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python profiles/time_workloads_fsea.py 4 > gpurun_out/final_workloads_fsea.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
tail -2 gpurun_out/final_gpu_tests.log; tail -1 gpurun_out/final_smoke.log
for f in headline config_1 config_2 config_3 config_4 config_5; do python - <<PY
import json
d=json.loads(open("gpurun_out/final_bench_$f.json").read().strip().splitlines()[-1])
print("$f", "%.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], "ms %.2f"%d["ms_per_step"], d["roofline"]["kernel"], "%.2f"%d["roofline"]["frac"], "cpu", d.get("cpu_baseline",{}).get("value"))
PY
done
tail -c 400 gpurun_out/final_reference_arm.json
