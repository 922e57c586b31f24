nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -DDHR_K2_TRACE -o gpurun_out/k2_micro tools/k2_micro.cu -lcuda > gpurun_out/k2_probe_b.txt 2>&1
for tau in 1e30 0.06; do
  echo "== k2_micro 221045 256 768 1 0 1 $tau" >> gpurun_out/k2_probe_b.txt
  timeout 120 ./gpurun_out/k2_micro 221045 256 768 1 0 1 $tau >> gpurun_out/k2_probe_b.txt 2>&1
done
head -40 gpurun_out/k2_probe_b.txt
