for v in base b1 b2 b8; do
  SD_B200_LIB=$PWD/scenedino_b200/build/variants/lib_$v.so timeout 60 python profiles/time_bin_k.py 2>&1 | tail -1 | cut -c1-200
done
