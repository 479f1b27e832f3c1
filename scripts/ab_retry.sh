run() { env "$@" python bench.py --workload $W --steps 2 --warmup 2 --no-cpu-baseline $X 2>&1 | grep "^{" | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('%.4e %.3f' % (j['value'], j['roofline']['frac']))"; }
for W in rober orego vanderpol; do X=""; [ $W = rober ] && X="--trajectories 262144"; echo "== $W new / old (NVRTC both)"; run PNODE_FORCE_NVRTC=1; run PNODE_FORCE_NVRTC=1 PNODE_NVRTC_FLAGS=-DPN_RETRY_IN_PHASE1; done
