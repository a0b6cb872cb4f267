#!/bin/bash
scripts/ab.sh run d0 d1 d2 f0 f1 f2 2>&1 | grep -v "round 1"
