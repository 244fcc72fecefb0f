#!/bin/bash
# A/B of two builds of libomr_b200.so on the stage microbench (scripts/stage_times.py): OMR_B200_LIB selects the library.
#   scripts/ab_stage.sh <variant.so> [batches...]     -> prints "A (variant)" and "B (in-tree)" lines, two rounds each
VAR=$1; shift
B=${*:-"2368 8192"}
for round in 1 2; do
  echo "== A $VAR"; OMR_B200_LIB=$VAR python scripts/stage_times.py --batch $B --reps 2
  echo "== B in-tree"; python scripts/stage_times.py --batch $B --reps 2
done
