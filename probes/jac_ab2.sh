for cfg in "8000000 512" "2000000 512" "1000000 992"; do
  timeout 200 python probes/r2_reps.py $cfg 2>&1 | tail -1
  PL_JACOBI_MULTILAUNCH=1 timeout 200 python probes/r2_reps.py $cfg 2>&1 | tail -1
  PL_NO_SVD_OVERLAP=1 timeout 200 python probes/r2_reps.py $cfg 2>&1 | tail -1
done
timeout 200 python probes/r2_reps.py 8000000 512 2>&1 | tail -1
PL_JACOBI_MULTILAUNCH=1 timeout 200 python probes/r2_reps.py 8000000 512 2>&1 | tail -1
