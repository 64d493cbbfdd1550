set -x
CMD="python bench.py --segments 74 --points 4096 --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo launches_exit=$?
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 33 -c 2 -f -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu2.log 2>&1
echo full_exit=$?
tail -3 gpurun_out/plain.log; tail -5 gpurun_out/ncu2.log
