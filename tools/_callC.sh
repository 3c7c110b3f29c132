set -x
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/C_test_main.log
B2F_LIB=libflate_b200/libb2f_fr.so python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/C_test_fr.log
python tools/stage_times.py 265 A > gpurun_out/C_stage_main.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spec_resolve|k_spec_round|k_validate|k_find_blocks" -c 16 -f -o gpurun_out/prof_C python tools/stage_times.py 64 A > gpurun_out/C_ncu.log 2>&1
cat gpurun_out/C_test_*.log gpurun_out/C_stage_*.log
