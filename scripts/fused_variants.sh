#!/bin/bash
# Runs scripts/fused_quick.py once per library variant in egonerf_b200/variants/ (+ the main build), 2 rounds, for A/B timing.
for round in 1 2; do
  python scripts/fused_quick.py main_r$round 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['stage_ms'][0], d['stage_ms'][1], d['train_fwd_bwd_ms_16384'], d['rgb_linf_vs_parity_mode'])"
  for so in egonerf_b200/variants/libegn_*.so; do
    n=$(basename $so .so); n=${n#libegn_}
    EGN_B200_LIB=$PWD/$so python scripts/fused_quick.py ${n}_r$round 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['stage_ms'][0], d['stage_ms'][1], d['train_fwd_bwd_ms_16384'], d['rgb_linf_vs_parity_mode'])"
  done
done
