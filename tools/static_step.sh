#!/bin/bash
# Static schedule of the simulator-step basic block of the two headline instantiations (colav_iw, rl) of the fast build,
# from a two-instantiation compile (seconds): tools/static_step.sh [-DSENV_... flags]
cd "$(dirname "$0")/../ast_sac_b200/csrc" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I../../include -fmad=true -DSENV_ONLY_ONE "$@" \
  -Xptxas -v -c -o /tmp/static_step.o kernels_fast.cu 2> /tmp/static_step.log || { grep -i error /tmp/static_step.log; exit 1; }
grep -E "spill" /tmp/static_step.log | tr '\n' ' '; echo
python ../../tools/sass_sched.py /tmp/static_step.o 'k_envILi0ELi1ELi0ELi0E' --min 200 | tail -n +2
python ../../tools/sass_sched.py /tmp/static_step.o 'k_envILi1ELi2ELi0ELi0E' --min 200 | tail -n +2
