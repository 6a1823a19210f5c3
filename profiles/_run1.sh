for v in ahead0 prod ahead0 prod; do
echo "== $v mode 4"; SD_B200_LIB=$PWD/scenedino_b200/build/variants/lib_$v.so timeout 120 python profiles/stress_bin.py 100000 4 2>&1 | grep "FAILED\|done" | head -3
done
