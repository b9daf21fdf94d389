#!/bin/bash
set -u
mkdir -p gpurun_out
cat > /tmp/cvrun.py <<'PY'
import sys; sys.path.insert(0,'.')
import torch, rcppml_b200 as rb
eng=rb.Engine(0); eng.set_matrix_synthetic_sharded(1000000,100000,1e-3,20260101); eng.init_factors(64,42,0)
res,cv=eng.fit_cv(rb.make_config(64,max_iter=2,tol=0.0,solver_mode=1),holdout_fraction=0.1,cv_seed=7,mask_zeros=True)
print(res)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cv_half_step -s 2 -c 2 -o gpurun_out/prof_cv -f python /tmp/cvrun.py > gpurun_out/ncu_cv.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_cv.log
