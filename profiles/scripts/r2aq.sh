timeout 900 python -m pytest tests -m gpu -x -q -k "pretok or tokeniz or classes or small_call or per_line or config1 or punctuation or train_from_corpus or device_train_types" > gpurun_out/r2aq_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2aq_tests.log
timeout 300 python profiles/prof_pretok.py 1000000000 2>&1 | tail -3
