# development aid (round 2, second session): final validation + measurement pass on one GPU.  The A/B scripts of the
# session (balanced launch / lean step / quad kernel / GAE tile kernel / render_f32 launch modes) were variations of this
# file run through `gpurun`; their outputs are profiles/r02_quad_vs_warp*.txt, r02_gae_1gpu*.json, r02_render_f32_small_ab.txt.
python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_steps20.json 2> gpurun_out/r02_bench_1gpu_steps20.err; tail -2 gpurun_out/r02_bench_1gpu_steps20.err
python bench.py > gpurun_out/r02_bench_1gpu_default.json 2> gpurun_out/r02_bench_1gpu_default.err; tail -2 gpurun_out/r02_bench_1gpu_default.err
python tools/sweep.py --out gpurun_out/r02_sweep_1gpu.json --sizes 4096,8192,16384,24576,65536,262144,1048576 --compact 2>&1 | grep N=
python tools/bench_gae.py --out gpurun_out/r02_gae_1gpu.json 2>&1 | cut -c1-230 | tail -12
python tools/bench_render.py --out gpurun_out/r02_render_paths_1gpu.json 2>&1 | cut -c1-230
echo done
