#!/bin/bash
# Round-2 ncu captures (run on the GPU box: gpurun -- 'bash profiles/ncu_r2.sh'); summaries land in gpurun_out/ and are copied
# to profiles/r2/ by hand.  Numbers printed by runs under ncu are never bench values.
set -x
O=gpurun_out
# (1) launch list of the bench command itself (short run): every kernel with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --iters 12 --no-cpu --no-slab > $O/ncu_bench_stdout.log 2>&1
# (2) full set of the two headline half-step kernels at 300^3 (one launch each, after warm-up)
ncu --set full --import-source on --clock-control none -k regex:k_update_tma -s 20 -c 2 -o $O/ncu_tma300 python profiles/quick_bench.py 300 24 > /dev/null 2>&1
ncu -i $O/ncu_tma300.ncu-rep --page raw --csv > $O/ncu_tma300_raw.csv
ncu -i $O/ncu_tma300.ncu-rep --page details --csv > $O/ncu_tma300_details.csv
# (3) full set of the dispersive E half-step on the TMA kernel (soil tiled to 300^3): the second k_update_tma launch of an iteration
ncu --set full --import-source on --clock-control none -k regex:k_update_tma -s 21 -c 1 -o $O/ncu_disp300 python profiles/disp_bench.py 2 2 3 24 > /dev/null 2>&1
ncu -i $O/ncu_disp300.ncu-rep --page raw --csv > $O/ncu_disp300_raw.csv
ncu -i $O/ncu_disp300.ncu-rep --page details --csv > $O/ncu_disp300_details.csv
# (linked shards cannot be captured: ncu serialises the kernels of all streams, so a flag wait never sees its neighbour's
#  signal and runs into its time-out -- tried once, 34 waits x 20 s)
rm -f $O/ncu_tma300.ncu-rep $O/ncu_disp300.ncu-rep
python profiles/ncu_traffic.py $O/ncu_tma300_raw.csv $O/ncu_disp300_raw.csv
