#!/usr/bin/env bash
# One gpurun call: GPU parity tests, the default bench line, the ncu launch list and ONE `--set full` capture of every
# kernel of a step.  Usage (from the repo root, under gpurun):
#     gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r2u'
# The .ncu-rep files stay on the box (gpurun only brings back 64 MiB): their digests are made there -
#   <tag>_pytest.log, <tag>_bench.json, <tag>_launches.csv, <tag>_launch_shares.txt, <tag>_ncu_all_kernels.md,
#   <tag>_ncu_full_summary.txt, <tag>_traffic.json
# land in gpurun_out/; copy the ones worth keeping into profiles/.  SKIP_NCU=1: tests + bench only.
# KEEP_REPS="k_raster_fwd k_raster_bwd": additionally bring back --import-source reports of those kernels (7 MB each).
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1

echo "== pytest -m gpu"
python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log

echo "== bench (default flags)"
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; cut -c1-400 $OUT/${TAG}_bench.json

if [ "${SKIP_NCU:-0}" != "1" ]; then
  SHORT="python bench.py --steps 1 --warmup 3 --views 2 --no-cpu-baseline"
  echo "== ncu launch list"
  $SHORT > /tmp/${TAG}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $SHORT > /tmp/${TAG}_ncu_launches.log 2>&1
  echo "launch list rc=$?"
  python tools/launch_shares.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launch_shares.txt
  echo "== ncu --set full, every kernel of one view (skip the first 150 launches: set-up and warm-up)"
  ncu --set full --clock-control none -k "regex:k_" -s 150 -c 48 -f -o /tmp/${TAG}_full_all $SHORT > /tmp/${TAG}_ncu_all.log 2>&1
  echo "full rc=$?"
  python tools/ncu_table.py /tmp/${TAG}_full_all.ncu-rep > $OUT/${TAG}_ncu_all_kernels.md
  python tools/ncu_summary.py /tmp/${TAG}_full_all.ncu-rep > $OUT/${TAG}_ncu_full_summary.txt
  python tools/ncu_summary.py --traffic $OUT/${TAG}_traffic.json /tmp/${TAG}_full_all.ncu-rep > /dev/null
  for k in ${KEEP_REPS:-}; do
    ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o $OUT/${TAG}_full_$k $SHORT > /tmp/${TAG}_ncu_$k.log 2>&1
    echo "source-level report of $k rc=$?"
  done
fi
ls -la $OUT | tail -12
