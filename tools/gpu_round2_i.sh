set -o pipefail
mkdir -p gpurun_out
python -m pytest tests/test_gpu_adapter.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/t_a.log; echo "rc=$?" >> gpurun_out/t_a.log
python -m pytest tests/test_gpu_loss_fusion.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -10 > gpurun_out/t_b.log; echo "rc=$?" >> gpurun_out/t_b.log
tail -30 gpurun_out/t_a.log; tail -4 gpurun_out/t_b.log
