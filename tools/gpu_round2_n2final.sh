mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/dist_check.py > gpurun_out/j_dist_check_n2.log 2>&1; echo "rc=$?" >> gpurun_out/j_dist_check_n2.log
grep -c OK gpurun_out/j_dist_check_n2.log; tail -1 gpurun_out/j_dist_check_n2.log
$TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/j_bench_n2.json 2> gpurun_out/j_bench_n2.err
python -c "
import json
s=open('gpurun_out/j_bench_n2.json').read(); d=json.loads(s[s.find('{\"metric'):]); print('n2:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'])
" || tail -5 gpurun_out/j_bench_n2.err
