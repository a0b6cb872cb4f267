#!/bin/bash
B200ADSB_LIB=$PWD/variants/lib_p1.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
scripts/ab.sh run p0 p1
