#!/bin/bash
# One multi-GPU box visit (gpurun --gpus N): the bench under torchrun at N ranks and the tile-sharded render (row N2).
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_round_multi.sh <tag> 2'
tag=${1:-multi}; n=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/${tag}_smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${tag}_smi.txt 2>&1
nproc >> gpurun_out/${tag}_smi.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29501 \
    bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err; tail -c 1200 gpurun_out/${tag}_bench_n$n.json
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2>> gpurun_out/${tag}_bench_n$n.err
timeout 200 python -m pytest tests/test_film.py -m gpu -x -q -s > gpurun_out/${tag}_pytest_film.log 2>&1; tail -3 gpurun_out/${tag}_pytest_film.log
for integ in directlighting pathtracing; do
  aa=16; [ $integ = directlighting ] && aa=4
  timeout 400 python tools/render_sharded.py --integrator $integ --width 1920 --height 1080 --aa $aa >> gpurun_out/${tag}_render.jsonl 2>> gpurun_out/${tag}_render.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29502 \
      tools/render_sharded.py --integrator $integ --width 1920 --height 1080 --aa $aa --compare >> gpurun_out/${tag}_render.jsonl 2>> gpurun_out/${tag}_render.err
done
cat gpurun_out/${tag}_render.jsonl
tail -5 gpurun_out/${tag}_render.err
