N=${1:-4}
set -x
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 8 --warmup 3) > gpurun_out/r2z_bench$N.json 2>gpurun_out/r2z_bench$N.err
grep -o "\"ms_per_step\": [0-9.]*\|state_crc[^,}]*\|\"ms_upload[^}]*" gpurun_out/r2z_bench$N.json
(timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_mpi_dropin.py -m gpu -q 2>&1 | tail -5) > gpurun_out/r2z_pytest_multi$N.log 2>&1; tail -3 gpurun_out/r2z_pytest_multi$N.log
