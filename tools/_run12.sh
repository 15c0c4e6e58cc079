S4="--config example4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --spp 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sp_hit_kernel|sp_shade_kernel" -s 22 -c 3 -f -o gpurun_out/r2b_split_ex4 python bench.py $S4 > /dev/null 2>&1; echo rc=$?
