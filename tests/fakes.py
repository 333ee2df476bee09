"""Test-only stand-ins for the CUDA devices, backed by the CPU oracle (oracle/oracle_device.py).  They let the CPU
suite exercise the host logic of env.py (mode relinking, masks, planner compatibility) on a machine without a GPU.
Never imported by the product."""
from oracle.oracle_device import OracleAbstractDevice, OracleSceneDevice  # noqa: F401
