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
python tools/big_bench.py > gpurun_out/${T}_big_bench.jsonl 2>&1
# the long-segment kernel on batches that fill the chip (the traceback pool is sized at a byte per cell: raise the scratch limit)
(DP_PAIRS=3840 DP_SCRATCH_GB=40 python tools/dp_bench.py 2048 | head -1; DP_PAIRS=960 DP_SCRATCH_GB=40 python tools/dp_bench.py 4096 | head -1
 DP_PAIRS=480 DP_SCRATCH_GB=80 python tools/dp_bench.py 8192 | head -1; DP_PAIRS=120 DP_SCRATCH_GB=80 python tools/dp_bench.py 16384 | head -1) \
    > gpurun_out/${T}_dp_bench_filled.jsonl 2>&1
python bench.py --workload cfg4 --steps 5 --no-dp-micro --no-variants --cpu-seconds 5 > gpurun_out/${T}_bench_cfg4.json 2>> gpurun_out/${T}_bench.err
python tools/fasta_e2e.py > gpurun_out/${T}_fasta_e2e.json 2>> gpurun_out/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-dp-micro --no-variants > gpurun_out/${T}_ncu_launch.log 2>&1
# full set of the hot kernels (second step: pools sized); the reports are summarised here and deleted
# (gpurun brings back at most 64 MiB)
ncu --set full --clock-control none --import-source on -k regex:'stage12|dp_thread|asm_' --launch-skip 5 -c 5 \
    -o gpurun_out/${T}_full -f python bench.py --steps 1 --warmup 1 --no-cpu --no-dp-micro --no-variants > gpurun_out/${T}_ncu_full.log 2>&1
python tools/ncu_summaries.py ${T} gpurun_out/${T}_launches.csv gpurun_out/${T}_full.ncu-rep \
    "python bench.py --steps 2 --warmup 1 --no-cpu --no-dp-micro --no-variants (launch list) / --steps 1 --warmup 1 -k regex:'stage12|dp_thread|asm_' --launch-skip 5 -c 5 (full set)" > gpurun_out/${T}_summaries.log 2>&1
python tools/ncu_lines.py gpurun_out/${T}_full.ncu-rep stage12 0.006 > gpurun_out/${T}_stage12_lines.txt 2>&1
python tools/ncu_lines.py gpurun_out/${T}_full.ncu-rep asm_copy 0.01 > gpurun_out/${T}_asm_copy_lines.txt 2>&1
cp profiles/${T}_launches_summary.csv profiles/${T}_ncu_full_summary.csv gpurun_out/ 2>/dev/null
rm -f gpurun_out/${T}_full.ncu-rep
DP_PAIRS=3814 ncu --set full --clock-control none --import-source on -k regex:'dp_wf16|dp_tb16' -s 12 -c 6 \
    -o gpurun_out/${T}_dp1024 -f python tools/dp_bench.py 1024 > gpurun_out/${T}_ncu_dp1024.log 2>&1
python tools/ncu_dp_summary.py ${T}_dp1024 gpurun_out/${T}_dp1024.ncu-rep "DP_PAIRS=3814 python tools/dp_bench.py 1024 (-k regex:'dp_wf16|dp_tb16' -s 12 -c 6)" > /dev/null 2>&1
cp profiles/${T}_dp1024_ncu_full_summary.csv gpurun_out/ 2>/dev/null
rm -f gpurun_out/${T}_dp1024.ncu-rep
# long-segment strip kernel (dp_wf16s_kernel) + its traceback at L = 2048, and the 1 Mb pair's launch list
DP_PAIRS=953 ncu --set full --clock-control none --import-source on -k regex:'dp_wf16s|dp_tb16' -s 3 -c 3 \
    -o gpurun_out/${T}_dp2048 -f python tools/dp_bench.py 2048 > gpurun_out/${T}_ncu_dp2048.log 2>&1
python tools/ncu_dp_summary.py ${T}_dp2048 gpurun_out/${T}_dp2048.ncu-rep "DP_PAIRS=953 python tools/dp_bench.py 2048 (-k regex:'dp_wf16s|dp_tb16' -s 3 -c 3)" > /dev/null 2>&1
python tools/ncu_lines.py gpurun_out/${T}_dp2048.ncu-rep dp_wf16s 0.01 > gpurun_out/${T}_dp_wf16s_lines.txt 2>&1
cp profiles/${T}_dp2048_ncu_full_summary.csv gpurun_out/ 2>/dev/null
rm -f gpurun_out/${T}_dp2048.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_cfg5_launches.csv \
    python tools/cfg5_one.py > gpurun_out/${T}_ncu_cfg5.log 2>&1
# medium-pair stage 1 (by-reference table) on a cfg4 sample
CFG4_PAIRS=1480 ncu --set full --clock-control none --import-source on -k regex:'stage12_small' --launch-skip 2 -c 1 \
    -o gpurun_out/${T}_cfg4 -f python tools/big_bench.py cfg4 > gpurun_out/${T}_ncu_cfg4.log 2>&1
python tools/ncu_lines.py gpurun_out/${T}_cfg4.ncu-rep stage12_small 0.008 > gpurun_out/${T}_cfg4_stage1_lines.txt 2>&1
ncu -i gpurun_out/${T}_cfg4.ncu-rep --page raw --csv > gpurun_out/${T}_cfg4_stage1_raw.csv 2>/dev/null
rm -f gpurun_out/${T}_cfg4.ncu-rep
tail -2 gpurun_out/${T}_tests.log; tail -c 900 gpurun_out/${T}_bench.json
