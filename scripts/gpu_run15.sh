mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -q -k "conv3x3 or wgrad or train_step or double_conv or bn_train" --tb=line -p no:cacheprovider > gpurun_out/t_conv15.log 2>&1; echo "== conv tests exit $?"; tail -n 3 gpurun_out/t_conv15.log
timeout 600 python scripts/conv_microbench.py --batch 64 --layers 0,1,2,15,16,17 --kinds fprop,dgrad > gpurun_out/micro_v6.log 2>&1; echo "micro exit $?"; cat gpurun_out/micro_v6.log | tail -10
