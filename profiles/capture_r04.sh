#!/bin/bash
# Session 4 captures: the shared-memory-resident Breakout kernels (k_stem_tower16 incl. the stride-2 tail, k_recurrent16 re-mapped).
# Same recipe as capture_r03.sh: plain run first, ncu afterwards, summaries made on the box, reports dropped (64 MiB limit).
set -u
OUT=gpurun_out
mkdir -p $OUT
run() { echo "== $*" >> $OUT/r04_capture.log; "$@" >> $OUT/r04_capture.log 2>&1; echo "rc=$?" >> $OUT/r04_capture.log; }
run python tests/profile_stem16.py
python tests/profile_timeline.py breakout 16384 20 seq > $OUT/r04_timeline_breakout.txt 2>&1
run ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r04_launches_breakout.raw.csv python tests/profile_timeline.py breakout 16384 20
run ncu --set full --import-source on --clock-control none -k regex:k_stem_tower16 --launch-skip 9 --launch-count 3 -f -o $OUT/r04_k_stem_tower16 python tests/profile_stem16.py
run ncu --set full --import-source on --clock-control none -k regex:k_recurrent16 --launch-skip 50 --launch-count 1 -f -o $OUT/r04_k_recurrent16 python tests/profile_timeline.py breakout 16384 20
S="python profiles/summarize.py"
$S launches $OUT/r04_launches_breakout.raw.csv $OUT/r04_launches_breakout.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python tests/profile_timeline.py breakout 16384 20   (16,384 games, 20 simulations per move)"
$S kernel $OUT/r04_k_stem_tower16.ncu-rep $OUT/r04_ncu_k_stem_tower16.csv "ncu --set full --clock-control none --import-source on -k regex:k_stem_tower16 --launch-skip 9 --launch-count 3 python tests/profile_stem16.py   (16,384 frames; launch0 = 48x24 pair rows, 2 blocks + the stride-2 convolution; launch1 = 24x24, 3 blocks; launch2 = 12x12, 3 blocks)"
$S kernel $OUT/r04_k_recurrent16.ncu-rep $OUT/r04_ncu_k_recurrent16.csv "ncu --set full --clock-control none --import-source on -k regex:k_recurrent16 --launch-skip 50 --launch-count 1 python tests/profile_timeline.py breakout 16384 20   (one simulation of 16,384 games)"
rm -f $OUT/*.ncu-rep $OUT/*.raw.csv
tail -n 4 $OUT/r04_capture.log
