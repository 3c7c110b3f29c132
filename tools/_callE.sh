for v in v0 v1 v2 v3; do
  echo "== $v"
  B2F_LIB=libflate_b200/libb2f_$v.so python tools/stage_times.py 265 A 2>&1 | grep "decode stages"
done
B2F_LIB=libflate_b200/libb2f_v1.so python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
B2F_LIB=libflate_b200/libb2f_v2.so python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
