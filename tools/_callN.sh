set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/N_test_main.log
cat gpurun_out/N_test_main.log
for v in _stress _fr; do
B2F_LIB=libflate_b200/libb2f$v.so timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/N_test$v.log
cat gpurun_out/N_test$v.log
done
for v in "" _h8 _h32; do
  B2F_LIB=libflate_b200/libb2f$v.so timeout -s KILL 120 python tools/stage_times.py 265 A > gpurun_out/N_stage$v.log 2>&1
  echo "== variant '$v'"; grep "encode stages" gpurun_out/N_stage$v.log
done
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"k_lz_find|k_lz_fixup" -c 3 -f -o gpurun_out/prof_N python tools/stage_times.py 64 A > gpurun_out/N_ncu.log 2>&1
tail -2 gpurun_out/N_ncu.log
