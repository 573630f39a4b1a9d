N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/bench_${N}gpu_ref.json 2> gpurun_out/bench_${N}gpu_ref.err; echo rc=$?
wc -l gpurun_out/bench_${N}gpu.json gpurun_out/bench_${N}gpu_ref.json
