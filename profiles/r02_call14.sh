#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02n_launches_registration.csv python profiles/prof_registration.py > gpurun_out/r02n_prof.log 2>&1
tail -1 gpurun_out/r02n_prof.log | cut -c1-300
