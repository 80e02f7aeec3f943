# N=2: pull-kernel block cap x piece count x piece weights
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run() { # cap pieces weights
  B200S_KNOBS=2=$1 B200S_PIECE_WEIGHTS=$3 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-gpu-baseline --pieces $2 > gpurun_out/w_bench_$1_$2_$3.json 2> gpurun_out/w_bench_$1_$2_$3.err
  python -c "
import json
s=open('gpurun_out/w_bench_$1_$2_$3.json').read(); d=json.loads(s[s.find('{\"metric'):]); print('cap $1 pieces $2 $3:', d['ms_per_step'], d['value'], 'between', d.get('between_calls_ms'))
"
}
run 74 4 even
run 74 8 even
run 32 8 even
run 148 8 even
run 74 16 even
run 74 8 dec
B200S_KNOBS=2=74 $TR tools/dist_timeline.py 8 > gpurun_out/w_timeline_n2.log 2>&1; grep -A34 "rank 0 \[reduce" gpurun_out/w_timeline_n2.log | head -40
