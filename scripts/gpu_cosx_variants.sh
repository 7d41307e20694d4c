#!/bin/bash
# Developer experiment: whole-call COSX time for kernel variants built into occ_b200/variants/lib_<name>.so (OCC_XC_B200_LIB override).
# A variant = one COSX translation unit recompiled with e.g. -DCOSX_M10=8 (resident-CTA cap of class (1,0), k_cosx_esp.cuh)
# and linked with the other objects of occ_b200/csrc/build/:
#   nvcc <Makefile flags> -DCOSX_M10=8 -c k_cosx_la1.cu -o build/var/k_cosx_la1_m10_8.o
#   nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_m10_8.so <all build/*.o except k_cosx_la1.o> build/var/k_cosx_la1_m10_8.o
for v in default m10_8 m10_6 m00_10 m00_8 m11_7 m11_5 m20_6 m20_4 default; do
  if [ $v = default ]; then unset OCC_XC_B200_LIB; else export OCC_XC_B200_LIB=$PWD/occ_b200/variants/lib_$v.so; fi
  echo -n "$v "; python scripts/gpu_cosx_debug.py 2>&1 | tail -1
done
