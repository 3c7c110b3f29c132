set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/O_test_main.log
cat gpurun_out/O_test_main.log
for v in "" _h32; do
  B2F_LIB=libflate_b200/libb2f$v.so timeout -s KILL 120 python tools/stage_times.py 265 A > gpurun_out/O_stage$v.log 2>&1
  echo "== variant '$v'"; grep "stages\|wall" gpurun_out/O_stage$v.log
done
timeout -s KILL 200 python bench.py --skip-cpu > gpurun_out/O_bench.json 2> gpurun_out/O_bench.err
cat gpurun_out/O_bench.json
