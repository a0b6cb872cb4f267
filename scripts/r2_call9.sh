#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
scripts/ab.sh run s0 s1 s2
