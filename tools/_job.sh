timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
SAN_TIMEOUT=300 bash tools/sanitize.sh 2>&1 | tail -12
bash tools/ncu_workloads.sh 2>&1 | tail -8
ls -la gpurun_out/*.ncu-rep | tail -8
