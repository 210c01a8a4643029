set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "potrf or lean or lookahead or fullsize or trajectory or predict or panel_width or graph or medium" > gpurun_out/r02_t15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t15.log
tail -4 gpurun_out/r02_t15.log
for i in 1 2 3; do
timeout 120 python tools/profile_step.py | grep -E "factor|^step"
DQGP_NO_EARLY_TRTRI=1 timeout 120 python tools/profile_step.py | grep -E "factor|^step"
done
timeout 120 python tools/factor_breakdown.py 2>&1 | tail -6
