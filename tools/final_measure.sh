#!/bin/bash
# Round-end measurement pass (run under gpurun): tests, bench, ncu launch list + full set, microbenchmarks.
# usage: tools/final_measure.sh [tag]   (outputs gpurun_out/<tag>_*)
set -x
T=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${T}_tests.log
python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
python tools/dp_bench.py 8 16 32 64 128 256 512 1024 2048 4096 8192 > gpurun_out/${T}_dp_bench.jsonl 2>&1
DBGA_NO_WF16=1 python tools/dp_bench.py 64 128 256 512 1024 > gpurun_out/${T}_dp_bench_32bit.jsonl 2>&1
python tools/big_bench.py > gpurun_out/${T}_big_bench.jsonl 2>&1
python bench.py --workload cfg4 --steps 5 --no-dp-micro --no-variants --cpu-seconds 5 > gpurun_out/${T}_bench_cfg4.json 2>> gpurun_out/${T}_bench.err
python tools/fasta_e2e.py > gpurun_out/${T}_fasta_e2e.json 2>> gpurun_out/${T}_bench.err
./tools/bin/pipe_rates > gpurun_out/${T}_pipe_rates.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-dp-micro > gpurun_out/${T}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'stage|dp_|asm_|unpack|clear' -c 32 \
    -o gpurun_out/${T}_full -f python bench.py --steps 1 --warmup 1 --no-cpu --no-dp-micro > gpurun_out/${T}_ncu_full.log 2>&1
# DP microbenchmark kernels: third run of the batch (pools sized, caches warm): launches 12.. of the packed kernels
DP_PAIRS=3814 ncu --set full --clock-control none --import-source on -k regex:'dp_wf16|dp_tb16' -s 12 -c 6 \
    -o gpurun_out/${T}_dp1024 -f python tools/dp_bench.py 1024 > gpurun_out/${T}_ncu_dp1024.log 2>&1
DP_PAIRS=15258 ncu --set full --clock-control none --import-source on -k regex:'dp_wf16|dp_tb16' -s 12 -c 6 \
    -o gpurun_out/${T}_dp512 -f python tools/dp_bench.py 512 > gpurun_out/${T}_ncu_dp512.log 2>&1
tail -2 gpurun_out/${T}_tests.log; tail -c 900 gpurun_out/${T}_bench.json
