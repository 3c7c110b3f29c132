mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
for n in 4 6 8; do echo "slices=$n"; B2F_ENC_SLICES=$n python tools/e2e_times.py 2>&1 | grep -E "^pinned|^pageable|encode stages" | head -4; done
B2F_ENC_SLICES=8 timeout -s KILL 600 python bench.py --skip-cpu 2>/dev/null | cut -c1-200
timeout -s KILL 600 python bench.py --skip-cpu 2>/dev/null | cut -c1-200
