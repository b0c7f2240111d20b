set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest (tree: pair exponents in both forward kernels; half-block forward default)"; python -m pytest tests -m gpu -x -q -s > gpurun_out/r2zc_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2zc_pytest.log
echo "== pytest with the whole-block forward"; GSB_FWD_HALF=0 python -m pytest tests -m gpu -x -q -k "stagewise or fused_render or segmented or baseline_sizes or train_steps or golden" > gpurun_out/r2zc_pytest_whole.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2zc_pytest_whole.log
SKIP_TESTS=1 STEPS=5 VARIANTS="GSB_LIB=build_variants/base.so;GSB_FWD_HALF=1;GSB_FWD_HALF=0;GSB_FWD_HALF=1;GSB_FWD_HALF=0" bash tools/sweep_res.sh 2>&1 | tee gpurun_out/r2zc_sweep.txt
