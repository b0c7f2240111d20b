set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SKIP_TESTS=1 STEPS=5 VARIANTS="GSB_X=cmin12;GSB_LIB=build_variants/cmin6.so;GSB_LIB=build_variants/cmin24.so;GSB_LIB=build_variants/cmin48.so;GSB_X=cmin12b" bash tools/sweep_res.sh 2>&1 | tee gpurun_out/r2zf_sweep.txt
