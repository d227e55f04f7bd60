mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all tests exit $?"; tail -30 gpurun_out/t_all.log
