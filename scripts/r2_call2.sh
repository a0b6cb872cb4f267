#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25
scripts/ab.sh run base pad padcrc all allr0 stop1 stop2
