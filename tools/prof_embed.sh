# usage: bash tools/prof_embed.sh <tag>   (ncu --set full of the stand-alone point loader + first layer at 4M points: bench.py's roofline_hbm kernel)
TAG=${1:-r02}
CMD="python bench.py --segments 1024 --points 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_embed_$TAG.log 2>&1 && \
ncu --set full --clock-control none -k "regex:point_embed" -s 30 -c 1 -f -o gpurun_out/prof_embed_$TAG $CMD > gpurun_out/ncu_embed_$TAG.log 2>&1
echo embed_exit=$?
