# ncu --set full of ONE kernel (regex $2) in a serialised 64 MiB pass; summary + per-line table as text
T=$1; K=$2; OBJ=${3:-encode_kernels}
mkdir -p gpurun_out
B2F_DECODE_PARTS=1 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -c ${NCU_C:-1} -f -o /tmp/prof_$T python tools/stage_times.py 64 P > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
python tools/ncu_summary.py /tmp/prof_$T.ncu-rep > gpurun_out/${T}_summary.txt
python tools/ncu_lines.py /tmp/prof_$T.ncu-rep "$K" libflate_b200/libb2f.so $OBJ 70 > gpurun_out/${T}_lines.txt 2>&1
cat gpurun_out/${T}_summary.txt
