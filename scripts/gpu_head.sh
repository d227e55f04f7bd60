timeout 300 python scripts/elem_microbench.py 64 2>&1 | grep -i "head\|softmax"
