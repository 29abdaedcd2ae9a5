#!/bin/bash
# A/B of fused-kernel tuning variants on one box: bash scripts/ab_variants.sh v1 v2 ...
for v in "$@"; do
  APK_B200_LIB=$PWD/athenapk_b200/lib/$v.so timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', '%.4g'%d['value'], {k:round(x,3) for k,x in d['task_ms'].items()})"
done
