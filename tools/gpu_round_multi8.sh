#!/bin/bash
# One 8-GPU box visit (gpurun --gpus 8, charged 8x): the host's concurrent copy ceiling, the bench at N = 8 (e2e with per-rank NUMA
# binding), and BASELINE configs[2] as written -- 1920x1080 PathTracer 64 spp, tile-sharded at 1 / 2 / 4 / 8 GPUs.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_round_multi8.sh <tag> 8'
tag=${1:-multi8}; n=${2:-8}
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=index,name,clocks.sm --format=csv; nvidia-smi topo -m; echo "nproc $(nproc)"; lscpu | grep -i "numa\|socket\|model name"; free -g | head -2; } > gpurun_out/${tag}_box.txt 2>&1
for k in $n; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2960$k tools/pcie_peak.py >> gpurun_out/${tag}_pcie.jsonl 2>> gpurun_out/${tag}.err
done
timeout 60 python tools/pcie_peak.py >> gpurun_out/${tag}_pcie.jsonl 2>> gpurun_out/${tag}.err
cat gpurun_out/${tag}_pcie.jsonl | cut -c1-250
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29501 \
    bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2>> gpurun_out/${tag}.err; tail -c 900 gpurun_out/${tag}_bench_n$n.json
# configs[2]: 64 spp path tracing, 1 M triangles, 1080p
timeout 300 python tools/render_sharded.py --integrator pathtracing --width 1920 --height 1080 --aa 64 >> gpurun_out/${tag}_config2.jsonl 2>> gpurun_out/${tag}.err
for k in $(echo 2 4 $n | tr " " "\n" | sort -nu); do
  [ $k -le $n ] && timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2951$k \
      tools/render_sharded.py --integrator pathtracing --width 1920 --height 1080 --aa 64 >> gpurun_out/${tag}_config2.jsonl 2>> gpurun_out/${tag}.err
done
# the same frame on the stock CPU kd-tree with every core of the box (the reference arm of configs[2])
timeout 300 python tools/render_sharded.py --accelerator yafaray-kdtree-original --backend gloo --integrator pathtracing --width 1920 --height 1080 --aa 64 >> gpurun_out/${tag}_config2.jsonl 2>> gpurun_out/${tag}.err
cut -c1-420 gpurun_out/${tag}_config2.jsonl
tail -5 gpurun_out/${tag}.err
