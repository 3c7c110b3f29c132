# times the encode stages with every variant build present (tools/build_variant.sh)
mkdir -p gpurun_out
python tools/lz_time.py 265 2>&1 | tail -1
for so in libflate_b200/libb2f_*.so; do B2F_LIB=$so timeout -s KILL 120 python tools/lz_time.py 265 2>&1 | tail -1; done
