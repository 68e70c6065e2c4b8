# development aid (round 2, session 2): warp-kernel lean step / balanced launch / quad kernel -- parity, A/B, ncu capture (1 GPU)
V=$PWD/ppo-2dgrid_b200/lib/variants
python -m pytest tests/test_gpu_parity.py tests/test_gpu_policy_step.py -x -q > gpurun_out/s2_pytest.log 2>&1; tail -3 gpurun_out/s2_pytest.log
for pass in 1 2; do
for v in r2lean qplain default; do
  echo "== pass $pass variant $v"
  if [ $v = default ]; then unset MERLIN_B200_LIB; else export MERLIN_B200_LIB=$V/lib_$v.so; fi
  python tools/sweep.py --compact --modes rgb --steps 512 --sizes 256,1024,2048,4096,8192,16384,24576 2>&1 | grep "N="
done; done 2>&1 | tee gpurun_out/s2_ab.txt
unset MERLIN_B200_LIB
ncu --set full --import-source on --clock-control none -k regex:env_kernel_quad -s 30 -c 2 -o gpurun_out/s2_quad_n4096 -f python tools/step_loop.py 4096 40 rgb 1024 > gpurun_out/s2_ncu.log 2>&1; tail -2 gpurun_out/s2_ncu.log
echo done
