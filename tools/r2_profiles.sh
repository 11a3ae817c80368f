#!/usr/bin/env bash
# round-2 profile captures (GPU box): everything that is summarised under profiles/*_r02.json.  The ncu reports are
# summarised on the box (tools/ncu_summary.py) and deleted: only the small JSON / CSV files travel back.
set -u
O=gpurun_out/prof_r02; mkdir -p $O
python -c "import bench; print(bench.kernel_fingerprint())" > $O/traffic_r02.fingerprint
ncu --set full --clock-control none -k 'regex:mttkrp_dmma_kernel|pair_gemm_kernel' -s 2 -c 2 -f -o $O/traffic_r02 python tools/ncu_target_cfg.py 2 1 3 > $O/traffic_r02.log 2>&1
python tools/ncu_summary.py traffic $O/traffic_r02.ncu-rep $O/traffic_r02.fingerprint $O/traffic_r02.json
python tools/ncu_summary.py full $O/traffic_r02.ncu-rep $O/contractions_c2_full.json
rm -f $O/traffic_r02.ncu-rep
# launch lists (shares of the step) of the headline config, its 8-way shard, config 1, the 8-way shard of config 4
for t in "2 1 5 c2" "2 8 5 c2s8" "1 1 10 c1" "4 8 2 c4s8"; do
  set -- $t
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_$4.csv python tools/ncu_target_cfg.py $1 $2 $3 > $O/ll_$4.log 2>&1
  python tools/ncu_summary.py list $O/launches_$4.csv $O/launches_$4.json > /dev/null
  rm -f $O/launches_$4.csv
done
# full captures: small kernels at config 2, all kernels of the 8-way shard of config 2, the contraction of config 4's shard
ncu --set full --clock-control none -k 'regex:model_update|pair_leaf|pair_partial_reduce|mttkrp_reduce|sched_kernel|move_kernel' -s 12 -c 9 -f -o $O/small_c2 python tools/ncu_target_cfg.py 2 1 4 > $O/small_c2.log 2>&1
python tools/ncu_summary.py full $O/small_c2.ncu-rep $O/small_kernels_c2_full.json; rm -f $O/small_c2.ncu-rep
ncu --set full --clock-control none -k 'regex:mttkrp_dmma|pair_gemm|model_update|pair_leaf|pair_partial_reduce|mttkrp_reduce' -s 9 -c 9 -f -o $O/shard_c2s8 python tools/ncu_target_cfg.py 2 8 4 > $O/shard_c2s8.log 2>&1
python tools/ncu_summary.py full $O/shard_c2s8.ncu-rep $O/shard_c2s8_full.json; rm -f $O/shard_c2s8.ncu-rep
ncu --set full --clock-control none -k 'regex:mttkrp_dmma' -s 2 -c 2 -f -o $O/shard_c4s8 python tools/ncu_target_cfg.py 4 8 2 > $O/shard_c4s8.log 2>&1
python tools/ncu_summary.py full $O/shard_c4s8.ncu-rep $O/shard_c4s8_full.json; rm -f $O/shard_c4s8.ncu-rep
du -sh $O; ls $O
