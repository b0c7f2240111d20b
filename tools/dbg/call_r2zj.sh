set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SKIP_TESTS=1 STEPS=6 VARIANTS="GSB_X=hi_hi;GSB_FRONT_PRIO=0;GSB_TAIL_PRIO=0;GSB_FRONT_PRIO=0 GSB_TAIL_PRIO=0;GSB_X=hi_hi2" bash tools/sweep_res.sh 2>&1 | tee gpurun_out/r2zj_sweep.txt
