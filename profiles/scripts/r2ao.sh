timeout 600 python -m pytest tests -m gpu -x -q -k "direct_path or split_count or smoke or golden" > gpurun_out/r2ao_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2ao_tests.log
for k in wp bpe; do timeout 200 python profiles/scripts/launches.py 200000000 2 $k memo_off=1 2>&1 | tail -1; done
