#!/bin/bash
# Round-end measurement pass (run under gpurun): tests, bench, ncu launch list + full set, microbenchmarks.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_tests.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err
python tools/dp_bench.py 8 16 32 64 128 256 512 1024 2048 4096 8192 > gpurun_out/final_dp_bench.jsonl 2>&1
python tools/big_bench.py > gpurun_out/final_big_bench.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/final_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'stage|dp_|assemble|pair_layout|scan|clear' -c 26 \
    -o gpurun_out/final_full -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/final_ncu_full.log 2>&1
tail -2 gpurun_out/final_tests.log; tail -c 600 gpurun_out/final_bench.json
