# usage: bash tools/prof_cmd.sh <tag> [skip] [count]   (ncu launch list + full capture of one chain and one fusion launch)
TAG=${1:-r02}; SKIP=${2:-4}; CNT=${3:-2}
CMD="python bench.py --segments 74 --points 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
echo launches_exit=$?
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:gemm_pair|chain_pair" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
echo full_exit=$?
tail -2 gpurun_out/plain_$TAG.log | cut -c1-600
