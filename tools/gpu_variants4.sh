mkdir -p gpurun_out
for so in libflate_b200/libb2f_*.so; do echo $so; B2F_LIB=$so timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -k "decode or foreign or member or truncated" 2>&1 | tail -1; B2F_LIB=$so timeout -s KILL 100 python tools/resident_times.py 2>&1 | grep -E "overlap=True|dec:" | head -2; done
