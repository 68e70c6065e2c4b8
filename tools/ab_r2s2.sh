# development aid (round 2, session 2): GAE tile kernel, quad-kernel variant tests (1 GPU)
python -m pytest tests/test_gpu_parity.py -x -q -k "gae" > gpurun_out/s2_pytest.log 2>&1; tail -3 gpurun_out/s2_pytest.log
python -m pytest tests/test_gpu_variants.py -x -q -k "quad" > gpurun_out/s2_pytest2.log 2>&1; tail -3 gpurun_out/s2_pytest2.log
python tools/bench_gae.py --out gpurun_out/r02_gae_1gpu.json 2>&1 | tail -12
echo done
