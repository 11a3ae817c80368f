#!/usr/bin/env bash
# DRAM traffic of the two tensor-sized contractions of BASELINE config 2 for bench.py's roofline.traffic.
#
#   on the GPU box (under gpurun):   tools/refresh_traffic.sh capture
#       -> gpurun_out/traffic_r02.ncu-rep  (ncu --set full of one mttkrp_dmma_kernel and one pair_gemm_kernel launch)
#          gpurun_out/traffic_r02.fingerprint  (bench.kernel_fingerprint() of the sources that were profiled)
#   here, after the report came back:  tools/refresh_traffic.sh summarise
#       -> profiles/traffic_r02.json, stamped with that fingerprint; bench.py refuses it once csrc/ changes.
set -euo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
case "${1:-}" in
capture)
  mkdir -p gpurun_out
  ncu --set full --clock-control none --import-source on -k 'regex:mttkrp_dmma_kernel|pair_gemm_kernel' -s 2 -c 2 \
      -f -o gpurun_out/traffic_r02 python tools/ncu_target_cfg.py 2 1 3 > gpurun_out/traffic_r02.log 2>&1
  python -c "import bench; print(bench.kernel_fingerprint())" > gpurun_out/traffic_r02.fingerprint
  ;;
summarise)
  python tools/ncu_summary.py traffic gpurun_out/traffic_r02.ncu-rep gpurun_out/traffic_r02.fingerprint profiles/traffic_r02.json
  ;;
*)
  echo "usage: $0 capture|summarise" >&2
  exit 1
  ;;
esac
