set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "svd or jacobi or pod or complex" 2>&1 | tail -5
for cfg in "1000000 512" "2000000 512" "4000000 512" "8000000 512" "2000000 1000" "1000000 992"; do
  timeout 200 python probes/r2_cfg2.py $cfg 2>&1 | head -1
  PL_JACOBI_MULTILAUNCH=1 timeout 200 python probes/r2_cfg2.py $cfg 2>&1 | head -1
done
PL_JACOBI_OLD=1 timeout 200 python probes/r2_cfg2.py 2000000 512 2>&1 | head -1
PL_JACOBI_OLD=1 timeout 200 python probes/r2_cfg2.py 1000000 512 2>&1 | head -1
PL_JACOBI_OLD=1 timeout 200 python probes/r2_cfg2.py 2000000 1000 2>&1 | head -1
timeout 100 python probes/prof_jac.py 2>&1 | tail -12
