# K2 micro-benchmark probes -> gpurun_out/r2_k2_probe.txt : dense-only mode with finite admission thresholds (pass rates like the real search)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -DDHR_K2_TRACE -o gpurun_out/k2_micro tools/k2_micro.cu -lcuda > gpurun_out/r2_k2_probe.txt 2>&1
for tau in 1e30 0.08 0.07 0.06 0.05 0.04; do
  echo "== k2_micro 221045 256 768 1 0 1 $tau" >> gpurun_out/r2_k2_probe.txt
  timeout 120 ./gpurun_out/k2_micro 221045 256 768 1 0 1 $tau >> gpurun_out/r2_k2_probe.txt 2>&1
done
grep -E "==|mode 0|rror" gpurun_out/r2_k2_probe.txt
