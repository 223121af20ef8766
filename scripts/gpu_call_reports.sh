#!/bin/bash
mkdir -p gpurun_out
python scripts/env_nee_report.py > gpurun_out/r02_env_nee.md 2> gpurun_out/r02_env_nee.err; echo "env rc $?"; cat gpurun_out/r02_env_nee.md
PTB200_BUILD_TIMING=1 python scripts/build_stability.py 6 > gpurun_out/r02_build_stability.md 2> gpurun_out/r02_build_stability.err; echo "build rc $?"; cat gpurun_out/r02_build_stability.md; tail -40 gpurun_out/r02_build_stability.err
