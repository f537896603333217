#!/bin/bash
# Round-2 entry point for the two GEMM variants that were written without GPU access (csrc/gemm.cu):
#   MVD_GEMM_2CTA=1   CTA-pair (cta_group::2) tiles         MVD_GEMM_SPLITK=1   split-K for under-filled launches
#   MVD_GEMM_WS=1     weight-stationary 128-wide tiles for K <= 320 (add MVD_GEGLU_TILE=128 to cover the GEGLU GEMMs)
# For each flag: kernel parity tests, the tile-width sweep (auto column picks the variant), the step bench.
# Run under gpurun with a timeout; each variant is independent.
set -u
cd "$(dirname "$0")/.."
echo "=== TMA load throughput vs ring depth / sharing"
timeout 60 ./profiles/micro/tma_bw --depth 2>&1 | tail -12
for flag in MVD_GEMM_2CTA MVD_GEMM_SPLITK MVD_GEMM_WS; do
  echo "=== $flag=1"
  env $flag=1 timeout 300 python -m pytest tests/test_tc_kernels_gpu.py -x -q -k "linear or conv" 2>&1 | tail -3
  env $flag=1 timeout 200 python profiles/gemm_bn_sweep.py 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8}' | tail -40
  env $flag=1 timeout 200 python bench.py 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'])"
done
echo "=== cross-view model mode"
MVD_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_model_gpu.py -x -q -k cross_view_reference_mode 2>&1 | tail -3
