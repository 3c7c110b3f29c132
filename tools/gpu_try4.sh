mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
python tools/resident_times.py 2>&1 | grep -v Warn
python tools/e2e_times.py 2>&1 | grep -E "^pinned|^pageable|encode stages" | head -4
