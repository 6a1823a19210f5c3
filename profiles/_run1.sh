for v in a0 a1 a2 a2b5; do
  SD_B200_LIB=$PWD/scenedino_b200/build/variants/lib_$v.so timeout 60 python profiles/time_bin_k.py 2>&1 | tail -1 | cut -c1-200
done
SD_B200_LIB=$PWD/scenedino_b200/build/variants/lib_a2.so python profiles/time_bin.py 2>&1 | tail -1
