# BASELINE configurations 1..5 at N = 1, 2, 4, 8 GPUs of one box (weak scaling: fixed K-blocks per GPU), plus the whole
# 400^3 grid of configuration 2 through run() (--scaling strong).  bash profiles/scale_configs.sh <tag> "<N list>"
tag=$1; ns=${2:-"1 2 4 8"}
for cfg in 1 2 3 4 5; do
  for n in $ns; do
    if [ $n -eq 1 ]; then
      python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/${tag}_cfg${cfg}_n${n}.json 2> gpurun_out/${tag}_cfg${cfg}_n${n}.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + cfg * 10 + n)) bench.py --gpus $n --config $cfg --steps 5 --warmup 3 > gpurun_out/${tag}_cfg${cfg}_n${n}.json 2> gpurun_out/${tag}_cfg${cfg}_n${n}.err
    fi
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_cfg${cfg}_n${n}.json").read().strip().splitlines()[-1])
    print("config ${cfg} N=${n}: %.4g %s  e2e %.4g  ms/step %.2f  roofline %s %.2f" % (d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"]))
except Exception as err:
    print("config ${cfg} N=${n}: FAILED", err)
PY
  done
done
for n in $ns; do
  if [ $n -eq 1 ]; then
    python bench.py --config 2 --scaling strong --steps 1 --warmup 1 > gpurun_out/${tag}_strong2_n${n}.json 2> gpurun_out/${tag}_strong2_n${n}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --config 2 --scaling strong --steps 1 --warmup 1 > gpurun_out/${tag}_strong2_n${n}.json 2> gpurun_out/${tag}_strong2_n${n}.err
  fi
  tail -c 600 gpurun_out/${tag}_strong2_n${n}.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('strong config 2 N=${n}: %.4g %s  ms %.1f' % (d['value'], d['unit'], d['ms_per_step']))
except Exception as e: print('strong N=${n} FAILED', e)"
done
