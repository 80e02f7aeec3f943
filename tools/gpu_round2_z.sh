# N=8: every reducer against single-GPU gradients, bench line of the default mode (P2P pull) and of the NVLS pull
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/dist_check.py > gpurun_out/z_dist_check_n8.log 2>&1; echo "rc=$?" >> gpurun_out/z_dist_check_n8.log
grep -c OK gpurun_out/z_dist_check_n8.log; tail -2 gpurun_out/z_dist_check_n8.log
run() { # tag env
  env $2 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/z_bench_n8_$1.json 2> gpurun_out/z_bench_n8_$1.err
  python -c "
import json
s=open('gpurun_out/z_bench_n8_$1.json').read(); d=json.loads(s[s.find('{\"metric'):]); print('$1:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'between', d.get('between_calls_ms'))
" || tail -5 gpurun_out/z_bench_n8_$1.err
}
run p2p B200S_SCATTER_PULL=p2p
run nvls B200S_SCATTER_PULL=nvls
run p2p_again B200S_SCATTER_PULL=p2p
