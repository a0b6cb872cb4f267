#!/bin/bash
scripts/ab.sh run a0 a1 a2 a4 a16 a32 a17 a55 a0mb8 a0p0 2>&1 | grep -v "round 1"
