mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/t_unet9.log 2>&1; echo "== unet tests exit $?"; grep -E "logits vs|worst gradient|passed|failed|^E  |Error" gpurun_out/t_unet9.log | head -30
