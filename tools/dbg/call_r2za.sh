set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== microbench"; ./tools/microbench/issue_model | tee gpurun_out/r2za_issue_model.txt
echo "== pytest (tree = v1)"; python -m pytest tests -m gpu -x -q -s > gpurun_out/r2za_pytest_v1.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2za_pytest_v1.log
for v in v1_tadd v1_sacc v1_fchunk16; do
  echo "== raster tests with $v"
  GSB_LIB=build_variants/$v.so python -m pytest tests -m gpu -x -q -k "fused_render or segmented or baseline_sizes or train_steps or golden or view_pipeline" > gpurun_out/r2za_pytest_$v.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2za_pytest_$v.log
done
SKIP_TESTS=1 STEPS=5 VARIANTS="GSB_LIB=build_variants/base.so;GSB_LIB=build_variants/v1.so;GSB_LIB=build_variants/v1_tadd.so;GSB_LIB=build_variants/v1_sacc.so;GSB_LIB=build_variants/v1_fchunk16.so;GSB_LIB=build_variants/base.so" bash tools/sweep_res.sh 2>&1 | tee gpurun_out/r2za_sweep.txt
