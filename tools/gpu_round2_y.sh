# N=2: P2P pull: load flavour x piece weights
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run() { # tag env pieces
  env $2 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-gpu-baseline --pieces $3 > gpurun_out/y_bench_$1.json 2> gpurun_out/y_bench_$1.err
  python -c "
import json
s=open('gpurun_out/y_bench_$1.json').read(); d=json.loads(s[s.find('{\"metric'):]); print('$1:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'between', d.get('between_calls_ms'))
" || tail -5 gpurun_out/y_bench_$1.err
}
run cg_even4 "B200S_KNOBS=3=0" 4
run sys_even4 "B200S_KNOBS=3=1" 4
run cg_dec4 "B200S_KNOBS=3=0 B200S_PIECE_WEIGHTS=dec" 4
run cg_dec6 "B200S_KNOBS=3=0 B200S_PIECE_WEIGHTS=dec" 6
run cg_dec4_cap296 "B200S_KNOBS=3=0,2=296 B200S_PIECE_WEIGHTS=dec" 4
B200S_PIECE_WEIGHTS=dec $TR tools/dist_check.py > gpurun_out/y_dist_check_n2.log 2>&1; echo "rc=$?" >> gpurun_out/y_dist_check_n2.log
grep -c OK gpurun_out/y_dist_check_n2.log; tail -2 gpurun_out/y_dist_check_n2.log
B200S_PIECE_WEIGHTS=dec $TR tools/dist_timeline.py 4 > gpurun_out/y_timeline_n2.log 2>&1; grep -A14 "rank 0 \[reduce" gpurun_out/y_timeline_n2.log | head -16
