#!/bin/bash
# Round-2 final ncu captures (run under gpurun; every command first exits 0 without ncu).
set -e
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu --cascade 86 > gpurun_out/plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --cascade 86 > gpurun_out/ncu_launches.log 2>&1
for cfg in "30 10000000" "86 1000000" "90 1000000" "100 1000000" "120 1000000" "177 1000000"; do
    set -- $cfg
    python tools/run_case.py $1 $2 2 > /dev/null 2>&1
    ncu --set full --import-source on --clock-control none -k regex:bcp_kernel -s 1 -c 1 -o gpurun_out/r02g_k$1 -f \
        python tools/run_case.py $1 $2 2 > gpurun_out/ncu_k$1.log 2>&1
done
ncu --set full --import-source on --clock-control none -k "regex:mt19937|fill_stream" -c 12 -o gpurun_out/r02g_fill -f \
    python tools/run_case.py 30 10000000 1 > gpurun_out/ncu_fill.log 2>&1
echo captured
