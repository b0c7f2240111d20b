set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest"; python -m pytest tests -m gpu -x -q -s > gpurun_out/r2ze_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2ze_pytest.log
SKIP_TESTS=1 STEPS=8 VARIANTS="GSB_FLAGS=0;GSB_FLAGS=32;GSB_FLAGS=0;GSB_FLAGS=32" bash tools/sweep_res.sh 2>&1 | tee gpurun_out/r2ze_sweep.txt
