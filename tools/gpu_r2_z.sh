#!/bin/bash
# round 2, call Z: racecheck of the coarse path where the resident solve is ONE CTA (block barriers only)
O=gpurun_out/r2; mkdir -p $O
SGB_PROFILE=1 timeout 100 compute-sanitizer --tool racecheck python tools/sanitize_run.py coarse 2 60 100 > $O/z_racecheck_coarse_one_cta.log 2>&1
echo "rc=$?"; grep -E "coarse .* ok|RACECHECK SUMMARY|four lanes" $O/z_racecheck_coarse_one_cta.log | cut -c1-200 | head
