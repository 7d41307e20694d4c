for v in default m10_8 m10_6 m00_10 m00_8 m11_7 m11_5 m20_6 m20_4 default; do
  if [ $v = default ]; then unset OCC_XC_B200_LIB; else export OCC_XC_B200_LIB=$PWD/occ_b200/variants/lib_$v.so; fi
  echo -n "$v "; python scripts/gpu_cosx_debug.py 2>&1 | tail -1
done
