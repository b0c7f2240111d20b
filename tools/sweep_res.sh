#!/usr/bin/env bash
# residency sweep of the persistent rasterisers (tuning aid): prints ms/step with the view pipeline on and off
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in ${SWEEP:-"12 16" "9 12"}; do
  set -- $cfg
  GSB_FWD_RES=$1 GSB_BWD_RES=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --debug-overlap 2> gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline_kernels']
print('fwd_res $1 bwd_res $2: ms/step %.3f serialized %.3f e2e %.3f' % (d['ms_per_step'], d['ms_per_step_serialized'], d['e2e']['ms_per_step']))
print('  serial: ' + ', '.join('%s %.3f' % (n, v['ms']) for n, v in k.items()))"
  grep "overlap on" gpurun_out/sweep.err
done
