#!/bin/bash
# chained dX / dW kernel: parity first (short timeouts: a hung mbarrier must not eat the box), then A/B
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_gcn.py -x -q -k "dxdw or gemm_tn" > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2o_tests.log
tail -5 gpurun_out/r2o_tests.log
if grep -q "passed" gpurun_out/r2o_tests.log && ! grep -q "failed\|rc=124" gpurun_out/r2o_tests.log; then
  timeout 120 python -m pytest tests/test_gpu_engine.py -x -q -k "chained" > gpurun_out/r2o_engine.log 2>&1; echo "engine rc=$?" >> gpurun_out/r2o_engine.log
  tail -3 gpurun_out/r2o_engine.log
  timeout 100 python tools/dxdw_bench.py > gpurun_out/r2o_knobs.log 2>&1; cat gpurun_out/r2o_knobs.log
  GD_FUSED_DXDW=0 GD_LIB_TAG=two timeout 120 python tools/epoch_ab.py 200 > gpurun_out/r2o_epoch.log 2>&1
  GD_FUSED_DXDW=1 GD_LIB_TAG=chained timeout 120 python tools/epoch_ab.py 200 >> gpurun_out/r2o_epoch.log 2>&1
  tail -4 gpurun_out/r2o_epoch.log
fi
