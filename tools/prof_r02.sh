# usage: bash tools/prof_r02.sh <tag>   (ncu --set full of the folded-query attention kernel, the train attention kernels and rows_linear)
TAG=${1:-r02m}
python tools/forward_profile.py 256 4096 > gpurun_out/plain_attn_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:ctx_attn" -s 6 -c 1 -f -o gpurun_out/prof_attn_$TAG python tools/forward_profile.py 256 4096 > gpurun_out/ncu_attn_$TAG.log 2>&1
echo attn_exit=$?
python tools/train_attn_bench.py 1024 1024 0.1 > gpurun_out/train_attn_$TAG.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:train_attn" -s 4 -c 2 -f -o gpurun_out/prof_train_attn_$TAG python tools/train_attn_bench.py 1024 1024 0.1 > gpurun_out/ncu_train_attn_$TAG.log 2>&1
echo train_attn_exit=$?
python tools/latency_profile.py 1 1024 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:rows_linear" -s 40 -c 3 -f -o gpurun_out/prof_rows_$TAG python tools/latency_profile.py 1 1024 > gpurun_out/ncu_rows_$TAG.log 2>&1
echo rows_exit=$?
cat gpurun_out/train_attn_$TAG.json
