# development aid (round 2): A/B runs behind profiles/r02_*_ab.txt
V=ppo-2dgrid_b200/lib/variants
b() { python bench.py --steps 40 --warmup 5 --skip-learners --skip-cpu-baseline --skip-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   value %.4e  ms/step %.4f  frac %.3f  restarted %d' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['envs_restarted_on_a_new_layout_in_timed_region_rank0']))"; }
echo "== bench env-steps: layout pool 65536 (default) vs 8192; plain vs evict_last pool loads"
for i in 1 2; do
echo "L=65536 plain"; b
echo "L=65536 evict_last"; MERLIN_B200_LIB=$PWD/$V/lib_evict_last.so b
echo "L=8192 plain"; b --layouts 8192
echo "L=8192 evict_last"; MERLIN_B200_LIB=$PWD/$V/lib_evict_last.so b --layouts 8192
done
echo done
