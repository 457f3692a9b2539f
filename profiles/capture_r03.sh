#!/bin/bash
# Profile captures of the final kernels of round 2, session 3 (branch-free whole-search kernel, 256-bit convolution
# epilogue, one-kernel FC training step).  Run on the GPU box through gpurun from the repo root; results land in
# gpurun_out/ and are summarised into profiles/ by profiles/summarize_r03.sh here.  Each ncu command runs only after the
# same program has exited 0 without ncu.  One GPU, serialised kernels: compare SHARES / counters, never absolute times.
set -u
OUT=gpurun_out
mkdir -p $OUT
run() { echo "== $*" >> $OUT/r03_capture.log; "$@" >> $OUT/r03_capture.log 2>&1; echo "rc=$?" >> $OUT/r03_capture.log; }

# ---- plain runs first
run python tests/tune_fused.py cartpole 303104 default
run python tests/profile_resnet.py connect4 16384 6
run python tests/profile_resnet.py gomoku 4096 6
run python tests/profile_trainer.py cartpole

# ---- launch lists
export _TUNE_CHILD=1
run ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r03_launches_cartpole.raw.csv python tests/tune_fused.py cartpole 303104
unset _TUNE_CHILD
for w in connect4 gomoku; do
  G=$( [ $w = gomoku ] && echo 4096 || echo 16384 )
  run ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -s 100 -c 300 --csv \
      --log-file $OUT/r03_launches_$w.raw.csv python tests/profile_resnet.py $w $G 6
done
run ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_fc_train -c 20 --csv --log-file $OUT/r03_launches_trainer.raw.csv python tests/profile_trainer.py cartpole

# ---- full captures of the dominant kernels
export _TUNE_CHILD=1
run ncu --set full --import-source on --clock-control none -k regex:k_search_fc -s 3 -c 1 -f -o $OUT/r03_k_search_fc_cartpole python tests/tune_fused.py cartpole 303104
unset _TUNE_CHILD
run ncu --set full --import-source on --clock-control none -k regex:k_conv_tc -s 40 -c 4 -f -o $OUT/r03_k_conv_tc_connect4 python tests/profile_resnet.py connect4 16384 6
run ncu --set full --import-source on --clock-control none -k regex:k_conv_tc -s 41 -c 2 -f -o $OUT/r03_k_conv_tc_gomoku python tests/profile_resnet.py gomoku 4096 6
run ncu --set full --import-source on --clock-control none -k regex:k_fc_train -s 6 -c 1 -f -o $OUT/r03_k_fc_train_cartpole python tests/profile_trainer.py cartpole

# ---- in-step timelines (CUPTI through torch.profiler: warm caches, graph replay), launch by launch for one simulation
for w in connect4 gomoku breakout; do python tests/profile_timeline.py $w $( [ $w = gomoku ] && echo 4096 || echo 16384 ) 20 seq > $OUT/r03_timeline_$w.txt 2>&1; done

# ---- summaries are made HERE (gpurun copies back at most 64 MiB and one report is 17-28 MB); the reports are dropped
S="python profiles/summarize.py"
$S launches $OUT/r03_launches_cartpole.raw.csv $OUT/r03_launches_cartpole.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python tests/tune_fused.py cartpole 303104   (3 warm-up moves + 7 searches + 5 moves of 303,104 games)"
for w in connect4 gomoku; do
  G=$( [ $w = gomoku ] && echo 4096 || echo 16384 )
  $S launches $OUT/r03_launches_$w.raw.csv $OUT/r03_launches_$w.csv "ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active... --clock-control none -s 100 -c 300 python tests/profile_resnet.py $w $G 6"
done
$S launches $OUT/r03_launches_trainer.raw.csv $OUT/r03_launches_trainer.csv "ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_fc_train -c 20 python tests/profile_trainer.py cartpole   (128 samples x 11 unrolled steps)"
$S kernel $OUT/r03_k_search_fc_cartpole.ncu-rep $OUT/r03_ncu_k_search_fc_cartpole.csv "ncu --set full --clock-control none --import-source on -k regex:k_search_fc -s 3 -c 1 python tests/tune_fused.py cartpole 303104   (default whole-search kernel: branch-free scores, head loop; 303,104 games x 50 simulations)"
$S kernel $OUT/r03_k_conv_tc_connect4.ncu-rep $OUT/r03_ncu_k_conv_tc_connect4.csv "ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 40 -c 4 python tests/profile_resnet.py connect4 16384 6   (four consecutive tower layers: with residual, plain, with residual, plain)"
$S kernel $OUT/r03_k_conv_tc_gomoku.ncu-rep $OUT/r03_ncu_k_conv_tc_gomoku.csv "ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 41 -c 2 python tests/profile_resnet.py gomoku 4096 6"
$S kernel $OUT/r03_k_fc_train_cartpole.ncu-rep $OUT/r03_ncu_k_fc_train_cartpole.csv "ncu --set full --clock-control none --import-source on -k regex:k_fc_train -s 6 -c 1 python tests/profile_trainer.py cartpole   (one-kernel training step, 128 samples x 11 unrolled steps)"
ncu -i $OUT/r03_k_search_fc_cartpole.ncu-rep --page source --csv > $OUT/r03_k_search_fc_source.csv 2>/dev/null
ncu -i $OUT/r03_k_conv_tc_connect4.ncu-rep --page raw --csv > $OUT/r03_k_conv_tc_connect4_raw.csv 2>/dev/null
rm -f $OUT/*.ncu-rep $OUT/*.raw.csv
tail -n 12 $OUT/r03_capture.log
du -sh $OUT
