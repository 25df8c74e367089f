#!/bin/bash
# round-2 GPU call C: whole GPU suite, bench (driver flags), launch list, ncu --set full of the variance kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/r2c_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2c_tests.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c_bench.log 2>&1
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2c_bench_ref.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-sort-hv --distance-metric crowding --steps 20 --warmup 5 > gpurun_out/r2c_bench_crowding.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sort-hv --e2e-steps 1 --e2e-warmup 0 > gpurun_out/r2c_launches_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_var_tc3 -s 3 -c 1 -o gpurun_out/r2c_gp_var_tc3 python bench.py --no-cpu-baseline --no-sort-hv --steps 2 --warmup 1 --e2e-steps 1 --e2e-warmup 0 > gpurun_out/r2c_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hv3_tree_kernel -c 1 -o gpurun_out/r2c_hv3_tree python -c "
import numpy as np
from dmosopt_b200 import _lib as L
rng=np.random.default_rng(0); x=np.abs(rng.standard_normal((65536,3))); F=x/np.linalg.norm(x,axis=1,keepdims=True)*(1+0.01*rng.random((65536,1)))
print(L.hypervolume(F, F.max(0)+0.1))
" > gpurun_out/r2c_ncu_hv.log 2>&1
tail -25 gpurun_out/r2c_tests.log
for f in gpurun_out/r2c_bench.log gpurun_out/r2c_bench_ref.log gpurun_out/r2c_bench_crowding.log; do echo "== $f"; tail -1 $f | cut -c1-2500; done
du -sh gpurun_out
