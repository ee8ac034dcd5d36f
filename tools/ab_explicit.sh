#!/bin/bash
# A/B of the explicit Euler kernel's forms on cuda:0: PD_EXPL_OPT / PD_EXPL_UNROLL (integer forms of
# the comparisons, pd_explicit.cuh) and POPDYN_SUM_FORM (correction term of the compensated sum,
# compiler.traced_sum).   usage: tools/ab_explicit.sh > gpurun_out/ab_explicit.txt
for cfg in "seir 1048576 1825" "rmit 4194304 100"; do
  for form in twosum ordered; do
    for def in "-DPD_EXPL_OPT=0" "-DPD_EXPL_OPT=3" "-DPD_EXPL_UNROLL=1" "-DPD_EXPL_UNROLL=2"; do
      echo -n "sum=$form $def  "
      POPDYN_SUM_FORM=$form POPDYN_DEFINE="$def" python tools/run_explicit_like.py explicit $cfg 2>&1 | tail -1
    done
  done
done
