#!/bin/bash
# A/B of library variants on one GPU: tools/gpu_ab.sh "lib1 lib2 ..." "kernel,kernel,..." [extra kbench args]
mkdir -p gpurun_out
for lib in $1; do
  EK_THERMO_LIB=$lib python tools/kbench.py --realistic --only $2 $3 2>&1 | tee -a gpurun_out/r02_ab.log
done
