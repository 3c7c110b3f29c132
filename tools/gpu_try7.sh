mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
python tools/resident_times.py 2>&1 | grep -E "overlap=True|dec:" | head -2
B2F_LIB=libflate_b200/libb2f_sb1024.so timeout -s KILL 100 python tools/resident_times.py 2>&1 | grep -E "overlap=True|dec:" | head -2
