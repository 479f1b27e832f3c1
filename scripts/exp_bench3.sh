#!/bin/bash
run() { echo "== $*"; env "$@" python bench.py --steps 2 --warmup 1 --trajectories 262144 --no-cpu-baseline $EXTRA 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('value %.4e  kernel_ms %.2f  frac %.3f' % (j['value'], j['config']['kernel_ms_per_step_device_events'], j['roofline']['frac']))
    elif 'rror' in l: print(l.strip()[:300])
"; }
EXTRA=""
run PNODE_FORCE_NVRTC=1
run PNODE_FORCE_NVRTC=1 PNODE_THREADS=384 PNODE_MIN_BLOCKS=1
run PNODE_FORCE_NVRTC=1 PNODE_THREADS=512 PNODE_MIN_BLOCKS=1
run PNODE_FORCE_NVRTC=1 PNODE_THREADS=128 PNODE_MIN_BLOCKS=3
EXTRA="--lanes 8"
run PNODE_FORCE_NVRTC=1
run PNODE_FORCE_NVRTC=1 PNODE_THREADS=384 PNODE_MIN_BLOCKS=1
run PNODE_FORCE_NVRTC=1 PNODE_THREADS=512 PNODE_MIN_BLOCKS=1
