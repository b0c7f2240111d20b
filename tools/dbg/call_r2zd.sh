set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest (tree: compact sparse tail in the backward)"; python -m pytest tests -m gpu -x -q -s > gpurun_out/r2zd_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2zd_pytest.log
SKIP_TESTS=1 STEPS=5 VARIANTS="GSB_LIB=build_variants/nocompact.so;GSB_X=compact;GSB_LIB=build_variants/nocompact.so;GSB_X=compact" bash tools/sweep_res.sh 2>&1 | tee gpurun_out/r2zd_sweep.txt
