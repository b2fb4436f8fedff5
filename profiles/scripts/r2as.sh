timeout 600 python profiles/class_e2e.py 2>&1 | tail -1
