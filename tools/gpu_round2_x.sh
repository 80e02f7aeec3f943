# N=2: P2P pull against the NVLS pull: equivalence, bench lines, timeline
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/dist_check.py > gpurun_out/x_dist_check_n2.log 2>&1; echo "rc=$?" >> gpurun_out/x_dist_check_n2.log
grep -c OK gpurun_out/x_dist_check_n2.log; grep "scatter_grads\|range \[" gpurun_out/x_dist_check_n2.log | head -6; tail -2 gpurun_out/x_dist_check_n2.log
run() { # tag env... pieces
  env $2 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-gpu-baseline --pieces $3 > gpurun_out/x_bench_$1.json 2> gpurun_out/x_bench_$1.err
  python -c "
import json
s=open('gpurun_out/x_bench_$1.json').read(); d=json.loads(s[s.find('{\"metric'):]); print('$1:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'between', d.get('between_calls_ms'))
" || tail -5 gpurun_out/x_bench_$1.err
}
run p2p_4 B200S_SCATTER_PULL=p2p 4
run p2p_8 B200S_SCATTER_PULL=p2p 8
run p2p_4_cap74 "B200S_SCATTER_PULL=p2p B200S_KNOBS=2=74" 4
run p2p_4_cap296 "B200S_SCATTER_PULL=p2p B200S_KNOBS=2=296" 4
run nvls_4 B200S_SCATTER_PULL=nvls 4
$TR tools/dist_timeline.py 4 > gpurun_out/x_timeline_n2.log 2>&1; grep -A20 "rank 0 \[reduce" gpurun_out/x_timeline_n2.log | head -24
