#!/bin/bash
# final check of a build: every GPU parity test + the driver's smoke()
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/final_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
