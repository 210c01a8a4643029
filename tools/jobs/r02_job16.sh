set -x
DQGP_POTRF_TRACE=1 timeout 120 python tools/factor_breakdown.py 8192 2>&1 | grep -E "potrf trace|step  *(0|10|20|30|40|44|48|52|56|60|62):|n=8192" | head -40
