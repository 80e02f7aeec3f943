# N=4: the world size no other run of the round covered: every reducer against single-GPU gradients + the default bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/dist_check.py > gpurun_out/n4_dist_check.log 2>&1; echo "rc=$?" >> gpurun_out/n4_dist_check.log
grep -c OK gpurun_out/n4_dist_check.log; tail -2 gpurun_out/n4_dist_check.log
$TR bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err
python -c "
import json
s=open('gpurun_out/n4_bench.json').read(); d=json.loads(s[s.find('{\"metric'):]); print('n4:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['clocks'])
" || tail -5 gpurun_out/n4_bench.err
$TR bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/n4_bench_ref.json 2> gpurun_out/n4_bench_ref.err; cut -c1-400 gpurun_out/n4_bench_ref.json
