# development aid (round 2, session 2): render_f32 launch modes at small / mid frame counts (1 GPU)
V=$PWD/ppo-2dgrid_b200/lib/variants
for v in rf_old default rf_32k; do
  echo "== variant $v"
  if [ $v = default ]; then unset MERLIN_B200_LIB; else export MERLIN_B200_LIB=$V/lib_$v.so; fi
  python tools/bench_render.py --out gpurun_out/r02_render_$v.json 2>&1 | grep "render_f32 blocked'" | cut -c1-260
done 2>&1 | tee gpurun_out/s2_render_ab.txt
unset MERLIN_B200_LIB
python -m pytest tests/test_gpu_parity.py -x -q -k "render" 2>&1 | tail -2
echo done
