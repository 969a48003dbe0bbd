#!/bin/bash
# kernel time at the C5 shape vs accumulator segment length (ITJL_TC_SEG), full kernel and tensor chain alone
for seg in ${@:-13 17 25 50}; do for d in 0 2; do ITJL_TC_SEG=$seg ITJL_ABC_TC=1 ITJL_ABC_DEBUG=$d timeout 120 python benchmarks/tc_trials.py child 50 414 11840 2>&1 | tail -1 | sed "s/^/seg=$seg debug=$d /"; done; done
