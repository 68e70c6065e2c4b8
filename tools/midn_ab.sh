#!/bin/bash
# development aid: small / medium batch sizes, kernel choice x observation path
for choice in 2 3; do for path in 0 2; do
  echo "== kernel choice $choice, observation path $path"
  MERLIN_KERNEL_CHOICE=$choice MERLIN_OBSERVATION_PATH=$path python tools/sweep.py --compact --modes rgb --steps 512 --sizes 4096,8192,16384,24576,32768,65536,131072 2>&1 | grep "N="
done; done
