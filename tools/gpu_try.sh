# parity tests of the encode path + variant timings
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -x -q -k "not decode" 2>&1 | tail -3
bash tools/gpu_variants.sh
