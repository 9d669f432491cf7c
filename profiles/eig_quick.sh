# eigensolver check + per-kernel times: bash profiles/eig_quick.sh <tag>
python profiles/eig_stress.py > gpurun_out/$1_eig_stress.log 2>&1; grep -c "method 0" gpurun_out/$1_eig_stress.log; awk '/method 0/ {print $NF, $(NF-2), $(NF-4)}' gpurun_out/$1_eig_stress.log | sort | uniq -c | sort -rn | head -5
python -m pytest tests -m gpu -x -q -k "eigh" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['stage_ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/$1_launches.csv python profiles/prof_driver.py --blocks 32 --scans 2 > /dev/null 2>&1; python profiles/summarize.py launches gpurun_out/$1_launches.csv | head -9
