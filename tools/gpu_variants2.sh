mkdir -p gpurun_out
python tools/lz_time.py 265 2>&1 | tail -1
python tools/e2e_times.py 2>&1 | grep -E "^pinned|encode stages" | head -2
for so in libflate_b200/libb2f_*.so; do B2F_LIB=$so timeout -s KILL 120 python tools/lz_time.py 265 2>&1 | tail -1; B2F_LIB=$so timeout -s KILL 120 python tools/e2e_times.py 2>&1 | grep -E "^pinned|encode stages" | head -2; done
