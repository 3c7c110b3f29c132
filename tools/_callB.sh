set -x
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/B_test_main.log
B2F_LIB=libflate_b200/libb2f_atoms.so python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/B_test_atoms.log
B2F_LIB=libflate_b200/libb2f_atomsfr.so python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/B_test_atomsfr.log
python tools/stage_times.py 265 A > gpurun_out/B_stage_main.log 2>&1
B2F_LIB=libflate_b200/libb2f_atoms.so python tools/stage_times.py 265 A > gpurun_out/B_stage_atoms.log 2>&1
B2F_LIB=libflate_b200/libb2f_atomsfr.so python tools/stage_times.py 64 A > gpurun_out/B_stage_atomsfr.log 2>&1
B2F_LIB=libflate_b200/libb2f_atoms.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spec_resolve|k_lz_chain" -c 4 -f -o gpurun_out/prof_B python tools/stage_times.py 64 A > gpurun_out/B_ncu.log 2>&1
cat gpurun_out/B_test_*.log gpurun_out/B_stage_*.log
