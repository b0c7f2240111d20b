set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest (tree: half-block forward + backward tweaks)"; python -m pytest tests -m gpu -x -q -s > gpurun_out/r2zb_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2zb_pytest.log
echo "== raster tests with v2_sacc"
GSB_LIB=build_variants/v2_sacc.so python -m pytest tests -m gpu -x -q -k "fused_render or segmented or baseline_sizes or train_steps or golden or half_block" > gpurun_out/r2zb_pytest_sacc.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2zb_pytest_sacc.log
SKIP_TESTS=1 STEPS=5 VARIANTS="GSB_LIB=build_variants/base.so;GSB_X=tree;GSB_FWD_HALF=0;GSB_FWDH_RES=20;GSB_FWDH_RES=16;GSB_LIB=build_variants/v2_sacc.so;GSB_X=tree2;GSB_LIB=build_variants/base.so" bash tools/sweep_res.sh 2>&1 | tee gpurun_out/r2zb_sweep.txt
