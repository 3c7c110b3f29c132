mkdir -p gpurun_out
python tools/resident_times.py 2>&1 | grep -E "overlap=True|dec:" | head -2
for so in libflate_b200/libb2f_*.so; do echo $so; B2F_LIB=$so timeout -s KILL 200 python tools/resident_times.py 2>&1 | grep -E "overlap=True|dec:|rror|ssert" | head -3; done
