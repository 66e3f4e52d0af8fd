python -X faulthandler -m pytest tests -x -v -m gpu > gpurun_out/pytest_full.log 2>&1
tail -5 gpurun_out/pytest_full.log
