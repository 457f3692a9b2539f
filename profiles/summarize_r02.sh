#!/bin/bash
# gpurun_out/r02_* (profiles/capture_r02.sh) -> tracked summaries under profiles/
set -u
cd "$(dirname "$0")/.."
S="python profiles/summarize.py"
$S launches gpurun_out/r02_launches_cartpole.raw.csv profiles/r02_launches_cartpole.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python tests/tune_fused.py cartpole 303104   (3 warm-up moves + 7 searches + 5 moves of 303,104 games)"
for w in connect4 gomoku breakout; do
  G=$( [ $w = gomoku ] && echo 4096 || echo 16384 )
  $S launches gpurun_out/r02_launches_$w.raw.csv profiles/r02_launches_$w.csv "ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active... --clock-control none -s 100 -c 300 python tests/profile_resnet.py $w $G 6"
done
$S kernel gpurun_out/r02_k_search_fc_cartpole.ncu-rep profiles/r02_ncu_k_search_fc_cartpole.csv "ncu --set full --clock-control none --import-source on -k regex:k_search_fc -s 3 -c 1 python tests/tune_fused.py cartpole 303104   (default whole-search kernel: FFMA2, root in registers, smem path, blocked store; 303,104 games x 50 simulations)"
$S kernel gpurun_out/r02_k_conv_tc_connect4.ncu-rep profiles/r02_ncu_k_conv_tc_connect4.csv "ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 41 -c 1 python tests/profile_resnet.py connect4 16384 6"
$S kernel gpurun_out/r02_k_conv_tc_gomoku.ncu-rep profiles/r02_ncu_k_conv_tc_gomoku.csv "ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 41 -c 1 python tests/profile_resnet.py gomoku 4096 6"
$S kernel gpurun_out/r02_k_recurrent16_breakout.ncu-rep profiles/r02_ncu_breakout.csv "ncu --set full --clock-control none --import-source on -k regex:k_recurrent16 -s 8 -c 1 python tests/profile_resnet.py breakout 16384 6   (one-kernel recurrent inference, 16,384 images)"
$S kernel gpurun_out/r02_tree_kernels_connect4.ncu-rep profiles/r02_ncu_tree_kernels_connect4.csv "ncu --set full --clock-control none -k regex:k_select|k_expand_backup|k_root_init -s 6 -c 3 python tests/profile_resnet.py connect4 16384 6   (16,384 games, A = 7; launches in stream order)"
$S kernel gpurun_out/r02_tree_kernels_gomoku.ncu-rep profiles/r02_ncu_tree_kernels_gomoku.csv "ncu --set full --clock-control none -k regex:k_select|k_expand_backup|k_root_init -s 6 -c 3 python tests/profile_resnet.py gomoku 4096 6   (4,096 games, A = 121)"
$S kernel gpurun_out/r02_k_minmax_connect4.ncu-rep profiles/r02_ncu_k_minmax_connect4.csv "ncu --set full --clock-control none -k regex:k_minmax -s 5 -c 1 python tests/profile_resnet.py connect4 16384 6"
$S kernel gpurun_out/r02_k_head_mma_connect4.ncu-rep profiles/r02_ncu_k_head_mma_connect4.csv "ncu --set full --clock-control none -k regex:k_head_mma -s 8 -c 1 python tests/profile_resnet.py connect4 16384 6"
for w in connect4 gomoku breakout; do cp gpurun_out/r02_timeline_$w.txt profiles/r02_timeline_$w.txt; done
cp gpurun_out/r2_ubench_umma2.txt profiles/r02_ubench_umma_cta_group2.txt
cp gpurun_out/r2_ubench_hmma.txt profiles/r02_ubench_hmma.txt
cp gpurun_out/r2_parity_search.json profiles/r02_parity_search.json
cp gpurun_out/r2_parity_bf16_layers.json profiles/r02_parity_bf16_layers.json
# SASS excerpts of the shipped library: tcgen05 / TMA / TMEM mnemonics of the convolution, HMMA / LDSM of the narrow tower, FFMA2 of the whole-search kernel
cuobjdump -sass muzero_hypermodel_b200/libmzb200.so > /tmp/mzb_sass.txt
{
  echo "# cuobjdump -sass muzero_hypermodel_b200/libmzb200.so (sm_100a) - mnemonic counts over the whole library"
  for m in UTCHMMA UTMALDG LDTM UTCBAR UTCCP HMMA.16816.F32.BF16 LDSM FFMA2 SYNCS; do printf "%-22s %s\n" $m "$(grep -c "$m" /tmp/mzb_sass.txt)"; done
  echo; echo "# k_conv_tc<64, false, 0, 8, true> (connect4 tower layer): first tcgen05 / TMA / TMEM instructions"
  awk '/Function : .*k_conv_tcILi64ELb0ELi0ELi8ELb1E/{p=1} p&&/Function :/&&!/k_conv_tcILi64ELb0ELi0ELi8ELb1E/{p=0} p' /tmp/mzb_sass.txt | grep -E "UTCHMMA|UTMALDG|LDTM|UTCBAR|SYNCS|UTCATOM|ELECT" | head -40
  echo; echo "# k_recurrent16: ldmatrix / mma.sync stream of one convolution"
  awk '/Function : .*k_recurrent16/{p=1;next} p&&/Function :/{p=0} p' /tmp/mzb_sass.txt | grep -E "LDSM|HMMA" | head -48
  echo; echo "# k_search_fc<cartpole>: packed fp32 FMAs of the dynamics layer"
  awk '/Function : .*k_search_fc.*Li4ELi8ELi2ELi10ELi0ELi16ELi16ELi16ELi16EEELi256ELb1ELb1ELi135E/{p=1;next} p&&/Function :/{p=0} p' /tmp/mzb_sass.txt | grep -E "FFMA2|LDS.128" | head -24
} > profiles/r02_sass_excerpts.txt
wc -l profiles/r02_*
