set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SKIP_TESTS=1 STEPS=8 VARIANTS="GSB_X=sets3;GSB_LIB=build_variants/sets2.so;GSB_LIB=build_variants/sets4.so;GSB_X=sets3b" bash tools/sweep_res.sh 2>&1 | tee gpurun_out/r2zo_sweep.txt
