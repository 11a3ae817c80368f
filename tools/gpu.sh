#!/usr/bin/env bash
# Rebuild everything that travels to the GPU box, then run a command there:  tools/gpu.sh [--timeout S] -- '<cmd>'
set -euo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
make -s -C cp-cals_b200
make -s -C oracle
[ -d /root/reference ] && [ ! -x oracle/_ref/cals_ref ] && oracle/build_ref.sh
[ -d /root/reference ] && tools/build_ref_compat.sh >/dev/null
exec /usr/local/graft/bin/gpurun "$@"
