for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/scale_${N}.err | tail -1 > gpurun_out/scale_${N}.json; tail -c 700 gpurun_out/scale_${N}.json; echo
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --mode sharded --steps 10 --warmup 3 2>>gpurun_out/scale_8.err | tail -1 > gpurun_out/scale_sharded_8.json; tail -c 600 gpurun_out/scale_sharded_8.json
