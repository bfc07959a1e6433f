for k in 8 64 256 512; do
  CUDA_LAUNCH_BLOCKING=1 timeout 120 python tools/bench_configs.py --only c3 --c3k $k --reps 3 > gpurun_out/c3_$k.log 2>&1; echo "k=$k rc=$?"; tail -2 gpurun_out/c3_$k.log | cut -c1-220
done
