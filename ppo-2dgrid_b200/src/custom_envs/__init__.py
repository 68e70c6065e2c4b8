"""MERLIN environments backed by the CUDA env kernels.  `import src.custom_envs.register` keeps working as the
registration entry point (reference src/scenario_creator/scenario_creator.py:8)."""
