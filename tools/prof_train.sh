CMD="python tools/train_step_bench.py 1024 1024 --encoder"
$CMD > gpurun_out/train_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/ncu_train.log 2>&1
echo exit=$?
