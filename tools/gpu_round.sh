#!/usr/bin/env bash
# One gpurun call: GPU parity tests, the default bench line, the ncu launch list and one `--set full`
# capture per hot kernel.  Usage (from the repo root, under gpurun):
#     gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r1d "k_raster_bwd k_raster_fwd k_os_pass:6"'
# Outputs land in gpurun_out/<tag>_*; copy the summaries worth keeping into profiles/.
set -u
TAG=${1:-run}
KERNELS=${2:-"k_raster_bwd k_raster_fwd k_os_pass:6"}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1

echo "== pytest -m gpu"
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log

echo "== bench (default flags)"
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; cut -c1-400 $OUT/${TAG}_bench.json

if [ "${SKIP_NCU:-0}" != "1" ]; then
  SHORT="python bench.py --steps 1 --warmup 3 --views 2 --no-cpu-baseline"
  echo "== ncu launch list"
  $SHORT > $OUT/${TAG}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $SHORT > $OUT/${TAG}_ncu_launches.log 2>&1
  echo "launch list rc=$?"
  # one report per hot kernel; -s skips that kernel's first (cold) launches
  for spec in $KERNELS; do
    k=${spec%%:*}; skip=${spec#*:}; [ "$skip" = "$spec" ] && skip=2
    cnt=2; [ "$k" = "k_os_pass" ] && cnt=6
    echo "== ncu --set full on $k (skip $skip, count $cnt)"
    ncu --set full --clock-control none --import-source on -k "regex:$k" -s $skip -c $cnt -f -o $OUT/${TAG}_full_$k $SHORT > $OUT/${TAG}_ncu_full_$k.log 2>&1
    echo "full $k rc=$?"; tail -1 $OUT/${TAG}_ncu_full_$k.log
  done
fi
ls -la $OUT | tail -12
