# final multi-GPU measurement of round 2: bench lines at N GPUs (the driver's command) + the phase timing of one short run
N=${1:-8}
set -x
nvidia-smi -L | wc -l
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 8 --warmup 3) > gpurun_out/r2z_bench$N.json 2>gpurun_out/r2z_bench$N.err
grep -o "\"ms_per_step\": [0-9.]*\|state_crc[^,}]*" gpurun_out/r2z_bench$N.json
(B200_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 4 --warmup 3 --no-e2e) > gpurun_out/r2z_timing$N.json 2>gpurun_out/r2z_timing$N.err
grep "timing: step" gpurun_out/r2z_timing$N.err | tail -2
grep "timing rank 0: sidm pass of 10000000" gpurun_out/r2z_timing$N.err | tail -1
