#!/bin/bash
# 64-meshblock pack (the 8-GPU share of 256^3) on one GPU: x3-chunk count A/B
for k in "$@"; do
  APK_B200_KCHUNKS=$k timeout 300 python bench.py --n 128 --steps 20 --warmup 5 --no-e2e --no-cpu --no-parity 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kchunks=$k', '%.4g'%d['value'], {k:round(x,4) for k,x in d['task_ms'].items()})"
done
