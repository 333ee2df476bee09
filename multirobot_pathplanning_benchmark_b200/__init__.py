"""B200-native collision / proximity backend for the multi-robot multi-goal planners of
vhartman/multirobot-pathplanning-benchmark.

Layout: csrc/ (hand-written sm_100a kernels + the C ABI of include/mrb200.h), scene.py /
scenes.py (host scene model and blob compiler), backend.py (tensor-in / tensor-out batch
API), env.py (the reference's BaseProblem interface on top of it).
"""
from .scene import Scene, Tf, CompiledScene, compile_blob, algorithmic_flops_per_config  # noqa: F401
from .scenes import SCENES  # noqa: F401

__all__ = ["Scene", "Tf", "CompiledScene", "compile_blob", "SCENES", "algorithmic_flops_per_config"]
