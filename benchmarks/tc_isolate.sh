#!/bin/bash
# Isolation timings of k_abc_acf_tc at the C5 shape (ITJL_ABC_DEBUG bits: 1 no MMAs, 2 no simulate after the
# first particle, 8 no drain, 16 drain loads only, 32 drain without the lag-table atomics)
for d in ${@:-0 1 2 3}; do ITJL_ABC_TC=1 ITJL_ABC_DEBUG=$d timeout 120 python benchmarks/tc_trials.py child 50 414 11840 2>&1 | tail -1 | sed "s/^/debug=$d /"; done
