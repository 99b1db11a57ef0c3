(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "audio_prompts_of_different" 2>&1 | tail -12) > gpurun_out/r2_t68.log 2>&1
