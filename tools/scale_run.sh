#!/bin/bash
# One 8-GPU box: plain-copy control and the bench at N = 1, 2, 4, 8 (weak) and 8 (strong).  Outputs in gpurun_out/.
tag=${1:-r02m}
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
if [ -z "$SKIP_CTL" ]; then
for n in 8 4 2 1; do
  run $n tools/pcie_ctl.py --reps 5 2> gpurun_out/${tag}_pcie_ctl_${n}gpu.err | grep '^{' > gpurun_out/${tag}_pcie_ctl_${n}gpu.jsonl
done
fi
for n in 8 4 2; do
  run $n bench.py --gpus $n --steps 5 --warmup 3 --no-dp-micro > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
done
run 8 bench.py --gpus 8 --steps 5 --warmup 3 --no-dp-micro --scaling strong > gpurun_out/${tag}_bench_8gpu_strong.json 2> gpurun_out/${tag}_bench_8gpu_strong.err
timeout 300 python bench.py --steps 5 --no-dp-micro > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
timeout 200 python tools/multi_bench.py > gpurun_out/${tag}_multi_bench.jsonl 2> gpurun_out/${tag}_multi_bench.err
ls -la gpurun_out/${tag}_*
