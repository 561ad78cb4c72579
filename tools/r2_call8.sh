#!/bin/bash
# round 2, call 8 (1 GPU): the GPU test files not re-run since the K0 change of call 7
set -u
O=gpurun_out/r2c8
mkdir -p $O
timeout 420 python -m pytest tests/test_zz_edge_cases.py tests/test_sharded_pipeline.py tests/test_zzz_fuse_variants.py tests/test_cell_sharded_replay.py tests/test_zz_mesh.py tests/test_host_adaptor.py -m gpu -q -x -p no:cacheprovider > $O/gpu_tests.log 2>&1; echo "rc=$?"; tail -4 $O/gpu_tests.log
