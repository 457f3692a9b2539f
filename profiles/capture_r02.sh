#!/bin/bash
# Round-2 profile captures (run on the GPU box through gpurun from the repo root; results land in gpurun_out/ and are
# summarised into profiles/ by profiles/summarize_r02.sh here).  Each ncu command runs only after the same program
# has exited 0 without ncu.  One GPU, serialised kernels: compare SHARES / counters, never absolute step times.
set -u
OUT=gpurun_out
mkdir -p $OUT
run() { echo "== $*" >> $OUT/r02_capture.log; "$@" >> $OUT/r02_capture.log 2>&1; echo "rc=$?" >> $OUT/r02_capture.log; }

# ---- plain runs first
run python tests/tune_fused.py cartpole 303104 default
for w in connect4 gomoku breakout; do run python tests/profile_resnet.py $w $( [ $w = gomoku ] && echo 4096 || echo 16384 ) 6; done

# ---- launch lists (gpu__time_duration per launch)
export _TUNE_CHILD=1
run ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02_launches_cartpole.raw.csv python tests/tune_fused.py cartpole 303104
unset _TUNE_CHILD
for w in connect4 gomoku breakout; do
  G=$( [ $w = gomoku ] && echo 4096 || echo 16384 )
  run ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -s 100 -c 300 --csv \
      --log-file $OUT/r02_launches_$w.raw.csv python tests/profile_resnet.py $w $G 6
done

# ---- full captures of the dominant kernels
export _TUNE_CHILD=1
run ncu --set full --import-source on --clock-control none -k regex:k_search_fc -s 3 -c 1 -f -o $OUT/r02_k_search_fc_cartpole python tests/tune_fused.py cartpole 303104
unset _TUNE_CHILD
run ncu --set full --import-source on --clock-control none -k regex:k_conv_tc -s 41 -c 1 -f -o $OUT/r02_k_conv_tc_connect4 python tests/profile_resnet.py connect4 16384 6
run ncu --set full --import-source on --clock-control none -k regex:k_conv_tc -s 41 -c 1 -f -o $OUT/r02_k_conv_tc_gomoku python tests/profile_resnet.py gomoku 4096 6
run ncu --set full --import-source on --clock-control none -k regex:k_recurrent16 -s 8 -c 1 -f -o $OUT/r02_k_recurrent16_breakout python tests/profile_resnet.py breakout 16384 6
run ncu --set full --import-source on --clock-control none -k "regex:k_select|k_expand_backup|k_root_init" -s 6 -c 3 -f -o $OUT/r02_tree_kernels_connect4 python tests/profile_resnet.py connect4 16384 6
run ncu --set full --import-source on --clock-control none -k "regex:k_select|k_expand_backup|k_root_init" -s 6 -c 3 -f -o $OUT/r02_tree_kernels_gomoku python tests/profile_resnet.py gomoku 4096 6
run ncu --set full --import-source on --clock-control none -k regex:k_minmax -s 5 -c 1 -f -o $OUT/r02_k_minmax_connect4 python tests/profile_resnet.py connect4 16384 6
run ncu --set full --import-source on --clock-control none -k regex:k_head_mma -s 8 -c 1 -f -o $OUT/r02_k_head_mma_connect4 python tests/profile_resnet.py connect4 16384 6

# ---- in-step timelines (CUPTI through torch.profiler: warm caches, graph replay)
for w in connect4 gomoku breakout; do python tests/profile_timeline.py $w > $OUT/r02_timeline_$w.txt 2>&1; done
tail -n 40 $OUT/r02_capture.log
