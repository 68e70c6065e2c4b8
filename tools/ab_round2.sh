# development aid (round 2): A/B runs behind profiles/r02_*_ab.txt
V=ppo-2dgrid_b200/lib/variants
sw() { python tools/sweep.py --compact --steps 256 "$@" 2>&1 | grep N=; }
echo "== tile kernel 1M rgb: r01 / current (x2)"
python .ab/r01/tools/sweep.py --compact --steps 256 --sizes 1048576 --modes rgb 2>&1 | grep N=
sw --sizes 1048576 --modes rgb,symbolic
python .ab/r01/tools/sweep.py --compact --steps 256 --sizes 1048576 --modes rgb 2>&1 | grep N=
sw --sizes 1048576 --modes rgb,symbolic
echo "== small batches (auto choice)"
sw --sizes 32,1024,2048,4096,16384 --modes rgb,symbolic
echo "== fill floor"
python tools/fill_floor.py 2>&1 | grep N=
echo "== fomaml"
python tools/profile_fomaml.py 2>&1 | tail -1
python tools/count_kernels.py --out gpurun_out/r02_kernels_per_step_after.json > gpurun_out/ck5.log 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_policy_step.py -q -m gpu -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -3 gpurun_out/bench_b.err
echo done
