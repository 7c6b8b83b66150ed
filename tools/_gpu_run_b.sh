mkdir -p gpurun_out
for cfg in "f32 4 2" "f32 8 1" "f64 2 2" "f64 4 1" "f32 2 2" "f32 4 1" "f64 2 1"; do
  timeout 300 python tools/_repro_stream.py $cfg 4800 2>&1 | tail -2
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/_repro_stream.py f32 4 2 4800 > gpurun_out/r02u_memcheck.log 2>&1
grep -v "^=========     at\|Host Frame\|^=========         " gpurun_out/r02u_memcheck.log | head -30
