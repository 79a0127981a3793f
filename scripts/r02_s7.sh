#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/prof_step.py 8 > gpurun_out/s7_prof_step.txt 2>&1; cat gpurun_out/s7_prof_step.txt | cut -c1-170
