S2="--config example2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sp_level -s 8 -c 2 -f -o gpurun_out/r2b_ex2_level python bench.py $S2 > /dev/null 2>&1; echo "ncu ex2 rc=$?"
S4="--config example4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --spp 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sp_level -s 12 -c 2 -f -o gpurun_out/r2b_ex4_level python bench.py $S4 > /dev/null 2>&1; echo "ncu ex4 rc=$?"
