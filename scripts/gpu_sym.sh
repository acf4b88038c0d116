cd /root/repo
N=${1:-2}
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "replicas" 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/test_symmetric_multi.py > gpurun_out/sym_$N.log 2>&1
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" gpurun_out/sym_$N.log | tail -25
