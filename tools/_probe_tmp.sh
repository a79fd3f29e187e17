mkdir -p gpurun_out
timeout 120 python tools/attn_probe.py wavlm > gpurun_out/r02_attn_probe5w.log 2>&1; echo "probe wavlm exit $?"; tail -8 gpurun_out/r02_attn_probe5w.log
timeout 120 python tools/attn_probe.py whisper > gpurun_out/r02_attn_probe5h.log 2>&1; echo "probe whisper exit $?"; tail -8 gpurun_out/r02_attn_probe5h.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k attention > gpurun_out/r02_k12.log 2>&1; echo "kernel tests exit $?"; tail -12 gpurun_out/r02_k12.log
