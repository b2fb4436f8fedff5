timeout 1200 python tests/golden/make_config5_models.py gpurun_out 2>&1 | tail -5
ls -la gpurun_out
