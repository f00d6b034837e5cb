#!/bin/bash
# round 2, call U (1 GPU): host phases of sgb_set_graph on the 1M-pose graph (SGB_PROFILE) + warm end-to-end steps
O=gpurun_out/r2; mkdir -p $O
SGB_PROFILE=1 timeout 300 python tools/e2e_prof.py > $O/u_e2e_prof.log 2>&1
grep -E "set_graph\]|build_structure\]" $O/u_e2e_prof.log | tail -24; tail -3 $O/u_e2e_prof.log
