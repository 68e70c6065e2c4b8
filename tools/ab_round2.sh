# development aid (round 2): the measurement pass behind profiles/r02_* (1 GPU)
python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_steps20.json 2> gpurun_out/r02_bench_1gpu_steps20.err; tail -2 gpurun_out/r02_bench_1gpu_steps20.err
python bench.py > gpurun_out/r02_bench_1gpu_default.json 2> gpurun_out/r02_bench_1gpu_default.err; tail -2 gpurun_out/r02_bench_1gpu_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null
B="python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-e2e --skip-learners"
$B > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps20.csv $B > gpurun_out/ncu_b.log 2>&1
S="python tools/step_loop.py 1048576 16 rgb 65536 stagger"
$S > gpurun_out/plain_s.log 2>&1 && ncu --set full --clock-control none --import-source on -f -k regex:env_kernel_tile -s 12 -c 1 -o gpurun_out/r02_tile_n1m $S > gpurun_out/ncu_s.log 2>&1
S2="python tools/step_loop.py 1048576 16 symbolic 65536 stagger"
$S2 > gpurun_out/plain_s2.log 2>&1 && ncu --set full --clock-control none --import-source on -f -k regex:env_kernel_sym -s 12 -c 1 -o gpurun_out/r02_sym_n1m_v2 $S2 > gpurun_out/ncu_s2.log 2>&1
python tools/sweep.py --out gpurun_out/r02_sweep_1gpu.json --compact 2>&1 | grep N=
python tools/bench_gae.py > gpurun_out/r02_gae.log 2>&1; tail -3 gpurun_out/r02_gae.log
echo done
