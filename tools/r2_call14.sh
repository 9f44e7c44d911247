#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sinkhorn or regroup or reencode or generate_codes or division" > gpurun_out/r2m_pytest_sk.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m_pytest_sk.log; tail -n 15 gpurun_out/r2m_pytest_sk.log
timeout 300 python tools/run_round.py c2_slice 1000000 5 2>&1 | tail -5 | cut -c1-260
RQB200_SK_VARIANT=1 RQB200_NO_MEMO=1 timeout 300 python tools/run_round.py c2_slice 1000000 3 2>&1 | tail -3 | cut -c1-260
timeout 300 python tools/time_driver.py > gpurun_out/r2m_time_driver.txt 2>&1; tail -n 6 gpurun_out/r2m_time_driver.txt | cut -c1-200
