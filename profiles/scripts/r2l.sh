timeout 300 python profiles/small_call_split.py 2>&1 | tail -2
