python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15 > gpurun_out/D_test_main.log
python tools/stage_times.py 265 A > gpurun_out/D_stage_main.log 2>&1
cat gpurun_out/D_test_*.log gpurun_out/D_stage_*.log
