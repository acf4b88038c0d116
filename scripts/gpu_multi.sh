cd /root/repo
nvidia-smi -L
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err
tail -c 1200 gpurun_out/bench_c2_n$N.json; tail -5 gpurun_out/bench_c2_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --impl reference 2>&1 | tail -c 400
