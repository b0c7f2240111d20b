#!/usr/bin/env bash
# tuning aid: GPU tests, then bench variants given as "ENV=.. ENV=.." strings separated by ';' in $VARIANTS
# prints ms/step with the view pipeline on / off and the per-stage times in both modes
mkdir -p gpurun_out
[ "${SKIP_TESTS:-0}" = "1" ] || python -m pytest tests -m gpu -x -q 2>&1 | tail -3
IFS=';' read -ra VARS <<< "${VARIANTS:- }"
for v in "${VARS[@]}"; do
  env $v python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --debug-overlap 2> gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline_kernels']
print('[$v] ms/step %.3f serialized %.3f e2e %.3f  clocks %s' % (d['ms_per_step'], d['ms_per_step_serialized'], d['e2e']['ms_per_step'], d.get('clocks')))
print('  serial: ' + ', '.join('%s %.3f' % (n, v['ms']) for n, v in k.items()))"
  grep "overlap on" gpurun_out/sweep.err
done
