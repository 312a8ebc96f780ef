#!/bin/bash
# The round's measurement set on the GPU box (one B200): bench lines, launch list, ncu --set full of the DP kernels (and, with
# "all", of the config-3 kernel and the traceback kernel).  Everything goes to gpurun_out/ (scratch); summaries are copied into
# profiles/ afterwards.
O=gpurun_out
python bench.py > $O/r02f_bench_n1.json 2> $O/r02f_bench_n1.err
python bench.py --config 3 > $O/r02f_bench_c3.json 2> $O/r02f_bench_c3.err
python bench.py --impl reference > $O/r02f_bench_ref.json 2> $O/r02f_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/r02f_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:extz_dp16_kernel -c 1 -f -o $O/r02f_dp16_c2 python bench.py --steps 1 --warmup 1 --no-cpu > $O/r02f_ncu_c2.log 2>&1
if [ "$1" = "all" ]; then
ncu --set full --import-source on --clock-control none -k regex:extz_dp16_kernel -c 1 -f -o $O/r02f_dp16_c3 python bench.py --config 3 --steps 1 --warmup 1 --no-cpu > $O/r02f_ncu_c3.log 2>&1
ncu --set full --clock-control none -k regex:extz_traceback -c 1 -f -o $O/r02f_tb_c2 python bench.py --steps 1 --warmup 1 --no-cpu > $O/r02f_ncu_tb.log 2>&1
fi
for f in $O/r02f_bench_n1.json $O/r02f_bench_c3.json $O/r02f_bench_ref.json; do tail -1 $f | cut -c1-400; done
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv
