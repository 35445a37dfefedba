"""fs -- B200-native drop-in for the `fs` package of takah29/2d-fluid-simulator.

Same public names as /root/reference/fs (FluidSimulator, MacSolver, CipMacSolver, PressureUpdater,
JacobiPressureUpdater, RedBlackSorPressureUpdater, VorticityConfinement, BoundaryCondition,
DoubleBuffer, advect_upwind, advect_kk_scheme); device buffers are PyTorch CUDA tensors and every
per-step kernel is hand-written CUDA for sm_100a in libfs2d.so (C ABI: include/fs2d.h).
There is no CPU fallback.
"""
__all__ = ["advection", "boundary_condition", "double_buffer", "fluid_simulator", "pressure_updater", "solver",
           "vorticity_confinement", "distributed"]
