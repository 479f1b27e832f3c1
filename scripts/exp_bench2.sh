#!/bin/bash
# A/B of pn_small variants through NVRTC flags (developer knob); prints value / kernel ms / roofline fraction
run() { echo "== $*"; env "$@" python bench.py --steps 2 --warmup 1 --trajectories 262144 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('value %.4e  kernel_ms %.2f  frac %.3f' % (j['value'], j['config']['kernel_ms_per_step_device_events'], j['roofline']['frac']))
    elif 'rror' in l: print(l.strip())
"; }
run PNODE_FORCE_NVRTC=0
run PNODE_FORCE_NVRTC=1 PNODE_NVRTC_FLAGS=-DPN_QR_LOWER_NOISE
run PNODE_FORCE_NVRTC=1 PNODE_NVRTC_FLAGS=-DPN_QR_JC=4
run PNODE_FORCE_NVRTC=1 PNODE_NVRTC_FLAGS=-DPN_QR_JC=6
run PNODE_FORCE_NVRTC=1 PNODE_NVRTC_FLAGS=-DPN_QR_JC=8
