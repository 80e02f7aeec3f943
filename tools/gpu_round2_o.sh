mkdir -p gpurun_out
for k in 0 8; do
python bench.py --steps 10 --warmup 3 --no-cpu --no-gpu-baseline --knob 2=$k > gpurun_out/bench_k$k.json 2> gpurun_out/bench_k$k.err
done
python tools/sweep.py C1 C4 > gpurun_out/sweep_k0.jsonl 2>/dev/null
python -c "
import json
for k in (0,8):
    d=json.load(open('gpurun_out/bench_k%d.json'%k)); print(k, d['ms_per_step'], {a:v['ms'] for a,v in d['stages'].items()})
"
cat gpurun_out/sweep_k0.jsonl | cut -c1-330
