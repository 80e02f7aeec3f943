# N=2: unrolled NVLS pull + decreasing pieces: equivalence, timeline, bench
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/dist_check.py > gpurun_out/v_dist_check_n2.log 2>&1; echo "rc=$?" >> gpurun_out/v_dist_check_n2.log
grep -c OK gpurun_out/v_dist_check_n2.log; tail -2 gpurun_out/v_dist_check_n2.log
$TR tools/dist_timeline.py 4 > gpurun_out/v_timeline_n2.log 2>&1; grep -A22 "rank 0 \[reduce" gpurun_out/v_timeline_n2.log
$TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/v_bench_n2.json 2> gpurun_out/v_bench_n2.err
python -c "
import json
s=open('gpurun_out/v_bench_n2.json').read(); d=json.loads(s[s.find('{\"metric'):]); print('n2', d['ms_per_step'], d['value'], d['e2e']['value'])
"
