#!/bin/bash
# round-2 GPU call AA: ncu --set full of ehvi_kernel, rank_chain_kernel (bench sets), one radix onesweep pass, hvm_pair_kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --no-cpu-baseline --no-sort-hv --steps 2 --warmup 1 --e2e-steps 1 --e2e-warmup 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ehvi_kernel -c 1 -o gpurun_out/r2aa_ehvi $B > gpurun_out/r2aa_1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rank_chain_kernel -s 2 -c 1 -o gpurun_out/r2aa_rank_chain $B > gpurun_out/r2aa_2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:DeviceRadixSortOnesweepKernel -s 40 -c 1 -o gpurun_out/r2aa_onesweep $B > gpurun_out/r2aa_3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hvm_pair_kernel -c 1 -o gpurun_out/r2aa_hvm_pair python -c "
import numpy as np
from dmosopt_b200 import _lib as L
rng=np.random.default_rng(0); x=rng.random((400,7)); P=x/np.linalg.norm(x,axis=1,keepdims=True)*(1+0.1*rng.random((400,1)))
print(L.hypervolume(P, np.full(7,1.2)))
" > gpurun_out/r2aa_4.log 2>&1
ls -la gpurun_out/r2aa_*.ncu-rep; tail -1 gpurun_out/r2aa_4.log
