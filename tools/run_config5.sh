#!/bin/bash
# BASELINE config 5 end to end on one box: train the two checkpoints (PPO on GPU 0 and FOMAML on GPU 1 at the same
# time when the box has more than one GPU), then the scale-generalisation sweep -- hard 16/24/32/48/64 x 100 unseen
# seeds x {PPO zero-shot, FOMAML zero-shot, FOMAML few-shot} -- with the (arm, size) jobs spread over all GPUs.
#     tools/run_config5.sh [n_gpus] [out_dir]        (defaults: all visible GPUs, gpurun_out)
set -e
cd "$(dirname "$0")/.."
N=${1:-$(python -c "import torch; print(torch.cuda.device_count())")}
OUT=${2:-gpurun_out}
mkdir -p "$OUT"
PPO_STEPS=${PPO_STEPS:-5000000}
FOMAML_ITERS=${FOMAML_ITERS:-300}
t0=$(date +%s.%N)
if [ "$N" -gt 1 ]; then
  CUDA_VISIBLE_DEVICES=1 python tools/train_fomaml.py --difficulty hard --iterations "$FOMAML_ITERS" --save "$OUT/r02_fomaml_hard.pth" \
      --out "$OUT/r02_fomaml_hard_train.json" > "$OUT/r02_fomaml_hard_train.log" 2>&1 &
  pid=$!
  CUDA_VISIBLE_DEVICES=0 python tools/train_ppo.py --difficulty hard --total-steps "$PPO_STEPS" --save "$OUT/r02_ppo_hard.pth" \
      --out "$OUT/r02_ppo_hard_train.json" > "$OUT/r02_ppo_hard_train.log" 2>&1
  wait $pid
else
  python tools/train_ppo.py --difficulty hard --total-steps "$PPO_STEPS" --save "$OUT/r02_ppo_hard.pth" \
      --out "$OUT/r02_ppo_hard_train.json" > "$OUT/r02_ppo_hard_train.log" 2>&1
  python tools/train_fomaml.py --difficulty hard --iterations "$FOMAML_ITERS" --save "$OUT/r02_fomaml_hard.pth" \
      --out "$OUT/r02_fomaml_hard_train.json" > "$OUT/r02_fomaml_hard_train.log" 2>&1
fi
t1=$(date +%s.%N)
echo "checkpoints trained in $(python -c "print(round($t1 - $t0, 1))") s"
SWEEP="tools/eval_sweep.py --ckpt $OUT/r02_ppo_hard.pth --fomaml-ckpt $OUT/r02_fomaml_hard.pth --difficulty hard \
  --sizes 16,24,32,48,64 --tasks 100 --adapt-steps 1 --out $OUT/r02_eval_sweep_${N}gpu.json"
if [ "$N" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 $SWEEP
else
  python $SWEEP
fi
t2=$(date +%s.%N)
echo "sweep (incl. process start-up) in $(python -c "print(round($t2 - $t1, 1))") s"
