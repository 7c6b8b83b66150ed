mkdir -p gpurun_out
B="python bench.py --workload c4d --length 8 --steps 1 --warmup 1 --no-cpu-baseline --configs none --c5-strong off --dtype f32 --stage-parts 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 40 -c 1 -f -o /tmp/r02x_pair $B --option tc_pair=on > gpurun_out/r02x_ncu1.log 2>&1
ncu -i /tmp/r02x_pair.ncu-rep --page raw --csv > gpurun_out/r02x_pair_raw.csv
ncu -i /tmp/r02x_pair.ncu-rep --page source --csv > gpurun_out/r02x_pair_source.csv 2>/dev/null
ncu -i /tmp/r02x_pair.ncu-rep --page details > gpurun_out/r02x_pair_details.txt
ls -la gpurun_out/ | tail
