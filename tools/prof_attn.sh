# usage: bash tools/prof_attn.sh <tag>   (ncu full capture of the folded-query attention kernel and of the scene kernels)
TAG=${1:-r01f}
CMD="python tools/forward_profile.py 256 4096"
$CMD > gpurun_out/plain_attn_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:ctx_attn" -s 6 -c 1 -f -o gpurun_out/prof_attn_$TAG $CMD > gpurun_out/ncu_attn_$TAG.log 2>&1
echo attn_exit=$?
CMD2="python tools/scene_bench.py 2000000 256 1024 0.3"
$CMD2 > gpurun_out/plain_scene_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:tube_crop|select_kernel|sample_keys" -s 6 -c 4 -f -o gpurun_out/prof_scene_$TAG $CMD2 > gpurun_out/ncu_scene_$TAG.log 2>&1
echo scene_exit=$?
tail -1 gpurun_out/plain_scene_$TAG.log | cut -c1-500
