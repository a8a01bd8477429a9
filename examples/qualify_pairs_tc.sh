#!/usr/bin/env bash
# Qualification run for making the tensor-core pair kernels (csrc/nais_pairs_tc.cu, csrc/nais_pairs_tc_bwd.cu) the default:
# the whole GPU suite and the C1 epoch with both opt-ins enabled, the C3 bench line (which times both modes), and one
# `ncu --set full` capture of each kernel.  Meant as ONE gpurun call:
#     gpurun --timeout 600 -- 'bash examples/qualify_pairs_tc.sh'
# Outputs land in gpurun_out/ (scratch); copy what is to be judged into profiles/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== GPU suite with NAIS_PAIRS_TC=1 NAIS_PAIRS_TC_BWD=1"
NAIS_PAIRS_TC=1 NAIS_PAIRS_TC_BWD=1 python -m pytest tests -m gpu -q > gpurun_out/q_suite_tc.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/q_suite_tc.log
echo "== C1 epoch + eval with both enabled"
NAIS_PAIRS_TC=1 NAIS_PAIRS_TC_BWD=1 timeout 300 python examples/c1_epoch_and_eval.py > gpurun_out/q_c1_tc.json 2> gpurun_out/q_c1_tc.err; echo "rc=$?"
echo "== C3 bench (fp32 step and tc_opt_in step in one line)"
python bench.py --mode train > gpurun_out/q_bench_train.json 2> gpurun_out/q_bench_train.err; echo "rc=$?"
echo "== ncu --set full of the two kernels"
NAIS_PAIRS_TC=1 NAIS_PAIRS_TC_BWD=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"pairs_fwd_tc|pairs_bwd_tc" -c 2 \
  -o gpurun_out/prof_pairs_tc_r2 python bench.py --mode train --steps 1 --warmup 1 > gpurun_out/q_ncu.log 2>&1; echo "rc=$?"
