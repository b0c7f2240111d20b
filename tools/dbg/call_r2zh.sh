set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest (tree: compact loop two Gaussians per trip)"; python -m pytest tests -m gpu -x -q > gpurun_out/r2zh_pytest.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2zh_pytest.log
SKIP_TESTS=1 STEPS=5 VARIANTS="GSB_LIB=build_variants/c1.so;GSB_X=c2;GSB_LIB=build_variants/c1.so;GSB_X=c2" bash tools/sweep_res.sh 2>&1 | tee gpurun_out/r2zh_sweep.txt
