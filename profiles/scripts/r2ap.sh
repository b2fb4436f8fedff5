timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ap_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2ap_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
