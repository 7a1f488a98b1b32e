set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 1200 --warmup 48 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_ref.json 2>> gpurun_out/bench_2gpu.err
GA3C_MAX_SECONDS=25 GA3C_GPU_NUM_WORLDS=65536 NGPU=2 MASTER_PORT=29535 timeout 300 ./train.sh TrainPhase1 > gpurun_out/train_2gpu.log 2>&1
cut -c1-400 gpurun_out/bench_2gpu.json; cut -c1-200 gpurun_out/bench_2gpu_ref.json; tail -6 gpurun_out/train_2gpu.log; tail -3 gpurun_out/bench_2gpu.err
