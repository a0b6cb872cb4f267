#!/bin/bash
# scripts/r2_ab.sh name...: A/B of variants/lib_<name>.so on one box (run under gpurun)
scripts/ab.sh run "$@"
