set -x
run() { # name, env...
  name=$1; shift
  (env "$@" B200_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e) > gpurun_out/r2n_$name.log 2>gpurun_out/r2n_$name.err
  grep "timing: step" gpurun_out/r2n_$name.err | tail -1
  grep -o "\"ms_per_step\": [0-9.]*\|state_crc[^,}]*" gpurun_out/r2n_$name.log
}
run early_o1 B200_OVERLAP=1
run early_o2 B200_OVERLAP=2
run late_o2 B200_OVERLAP=2 B200_LATE_GRAVITY_EXCHANGE=1
(timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "nccl" 2>&1 | tail -5) > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log
