#!/bin/bash
# Re-capture of the dominant kernels after the last changes of the round (L2 policies in the whole-search kernel, 8 waves =
# 606,208 games per launch as in bench.py; batch-norm scale folded into the convolution's weights).  Same recipe as
# profiles/capture_r03.sh: plain run first, ncu afterwards, summaries made on the box, reports dropped (64 MiB limit).
set -u
OUT=gpurun_out
mkdir -p $OUT
run() { echo "== $*" >> $OUT/r03f_capture.log; "$@" >> $OUT/r03f_capture.log 2>&1; echo "rc=$?" >> $OUT/r03f_capture.log; }
run python tests/tune_fused.py cartpole 606208 default
run python tests/profile_resnet.py connect4 16384 6
run python tests/profile_resnet.py gomoku 4096 6
export _TUNE_CHILD=1
run ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r03_launches_cartpole.raw.csv python tests/tune_fused.py cartpole 606208
run ncu --set full --import-source on --clock-control none -k regex:k_search_fc -s 3 -c 1 -f -o $OUT/r03_k_search_fc_cartpole python tests/tune_fused.py cartpole 606208
unset _TUNE_CHILD
run ncu --set full --import-source on --clock-control none -k regex:k_conv_tc -s 40 -c 4 -f -o $OUT/r03_k_conv_tc_connect4 python tests/profile_resnet.py connect4 16384 6
run ncu --set full --import-source on --clock-control none -k regex:k_conv_tc -s 41 -c 2 -f -o $OUT/r03_k_conv_tc_gomoku python tests/profile_resnet.py gomoku 4096 6
for w in connect4 gomoku breakout; do python tests/profile_timeline.py $w $( [ $w = gomoku ] && echo 4096 || echo 16384 ) 20 seq > $OUT/r03_timeline_$w.txt 2>&1; done
S="python profiles/summarize.py"
$S launches $OUT/r03_launches_cartpole.raw.csv $OUT/r03_launches_cartpole.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python tests/tune_fused.py cartpole 606208   (3 warm-up moves + 7 searches + 5 moves of 606,208 games)"
$S kernel $OUT/r03_k_search_fc_cartpole.ncu-rep $OUT/r03_ncu_k_search_fc_cartpole.csv "ncu --set full --clock-control none --import-source on -k regex:k_search_fc -s 3 -c 1 python tests/tune_fused.py cartpole 606208   (final whole-search kernel: branch-free scores, head loop, L2 policies; 606,208 games x 50 simulations = 8 waves)"
$S kernel $OUT/r03_k_conv_tc_connect4.ncu-rep $OUT/r03_ncu_k_conv_tc_connect4.csv "ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 40 -c 4 python tests/profile_resnet.py connect4 16384 6   (four consecutive tower layers: with residual, plain, with residual, plain; 256-bit epilogue accesses, scale folded into the weights)"
$S kernel $OUT/r03_k_conv_tc_gomoku.ncu-rep $OUT/r03_ncu_k_conv_tc_gomoku.csv "ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 41 -c 2 python tests/profile_resnet.py gomoku 4096 6   (plain layer, layer with residual)"
ncu -i $OUT/r03_k_search_fc_cartpole.ncu-rep --page source --csv > $OUT/r03_k_search_fc_source.csv 2>/dev/null
rm -f $OUT/*.ncu-rep $OUT/*.raw.csv
tail -n 6 $OUT/r03f_capture.log; du -sh $OUT
