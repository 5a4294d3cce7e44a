#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tube_link or cross_clip or label_agreement" > gpurun_out/r2p_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2p_tests.txt
tail -30 gpurun_out/r2p_tests.txt
