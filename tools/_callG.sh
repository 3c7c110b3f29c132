set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15 > gpurun_out/G_test_main.log
cat gpurun_out/G_test_main.log
timeout -s KILL 120 python tools/stage_times.py 265 A > gpurun_out/G_stage_main.log 2>&1
cat gpurun_out/G_stage_main.log
B2F_LIB=libflate_b200/libb2f_w32.so timeout -s KILL 120 python tools/stage_times.py 265 A > gpurun_out/G_stage_w32.log 2>&1
B2F_LIB=libflate_b200/libb2f_w8.so timeout -s KILL 120 python tools/stage_times.py 265 A > gpurun_out/G_stage_w8.log 2>&1
grep "encode stages" gpurun_out/G_stage_w32.log gpurun_out/G_stage_w8.log
B2F_LIB=libflate_b200/libb2f_fr.so timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/G_test_fr.log
cat gpurun_out/G_test_fr.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"k_lz_find|k_spec_resolve" -c 6 -f -o gpurun_out/prof_G python tools/stage_times.py 64 A > gpurun_out/G_ncu.log 2>&1
tail -2 gpurun_out/G_ncu.log
