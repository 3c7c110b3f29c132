set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/H_test_main.log
cat gpurun_out/H_test_main.log
for v in "" _spin _w32 _w8 _nowalk; do
  B2F_LIB=libflate_b200/libb2f$v.so timeout -s KILL 120 python tools/stage_times.py 265 A > gpurun_out/H_stage$v.log 2>&1
  echo "== variant '$v'"; grep "encode stages" gpurun_out/H_stage$v.log
done
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"k_lz_find|k_parse_exits" -c 4 -f -o gpurun_out/prof_H python tools/stage_times.py 64 A > gpurun_out/H_ncu.log 2>&1
tail -2 gpurun_out/H_ncu.log
