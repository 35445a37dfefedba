"""Headless driver with the reference's command line (/root/reference/main.py:10-51).

Same flags and defaults (-bc, -re, -res, -dt, -vis, -vc, -scheme, -no_dye); the GGUI window, key
handling and PNG screenshots are out of scope, so instead of an event loop this runs
`--steps` time steps on the GPU and can dump `{"v", "p"[, "dye"]}` to `output/step_%06d.npz`
(the `d`-key format, main.py:129-132).  Extras: `--jacobi N` selects the Jacobi updater used by the
BASELINE configs (the reference hard-codes RB-SOR 1.3 x2, fs/fluid_simulator.py:76-78).
`-cpu` is rejected: this build has no CPU path.
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))


def main() -> None:
    parser = argparse.ArgumentParser(description="Fluid Simulator (B200 build, headless)")
    parser.add_argument("-bc", "--boundary_condition", help="Boundary condition number", type=int,
                        choices=[1, 2, 3, 4, 5, 6], default=1)
    parser.add_argument("-re", "--reynolds_num", help="Reynolds number", type=float, default=1000000.0)
    parser.add_argument("-res", "--resolution", help="Resolution of y-axis", type=int, default=400)
    parser.add_argument("-dt", "--time_step", help="Time step", type=float, default=0.0)
    parser.add_argument("-vis", "--visualization", help="Flow visualization type (ignored: headless)", type=int,
                        choices=[0, 1, 2, 3], default=0)
    parser.add_argument("-vc", "--vorticity_confinement", help="Vorticity Confinement. 0.0 is disable.", type=float,
                        default=5.0)
    parser.add_argument("-scheme", "--advection_scheme", help="Advection Scheme", type=str,
                        choices=["upwind", "kk", "cip"], default="cip")
    parser.add_argument("-no_dye", "--no_dye", help="No dye calculation", action="store_true")
    parser.add_argument("-cpu", "--cpu", action="store_true")
    parser.add_argument("--steps", type=int, default=100, help="time steps to run (replaces the window loop)")
    parser.add_argument("--dump-every", type=int, default=0, help="write output/step_%%06d.npz every N steps")
    parser.add_argument("--output", type=str, default=str(Path(__file__).parent.resolve() / "output"))
    parser.add_argument("--jacobi", type=int, default=0, help="use JacobiPressureUpdater with N sweeps/step")
    parser.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying CUDA graphs")
    parser.add_argument("--obstacle-image", type=str, default=None,
                        help="-bc 6: path of the reference's images/bc_mask/dragon.png (an asset of the reference, not shipped here)")
    args = parser.parse_args()
    if args.boundary_condition == 6:
        default_png = Path(__file__).resolve().parent / "images" / "bc_mask" / "dragon.png"
        if args.obstacle_image is None and not default_png.exists():
            parser.error("-bc 6 needs the reference's obstacle image: pass --obstacle-image /path/to/images/bc_mask/dragon.png")

    if args.cpu:
        raise SystemExit("-cpu: this build runs on B200 only (hand-written sm_100a kernels, no CPU fallback)")

    import torch

    from fs.fluid_simulator import DyeFluidSimulator, FluidSimulator

    n_bc, re, resolution = args.boundary_condition, args.reynolds_num, args.resolution
    dt = args.time_step if args.time_step != 0.0 else 0.05 / resolution
    vor_eps = args.vorticity_confinement if args.vorticity_confinement != 0.0 else None
    dx = 1 / resolution
    scheme = args.advection_scheme
    print(f"Boundary Condition: {n_bc}\ndt: {dt}\nRe: {re}\nResolution: {resolution}\n"
          f"Scheme: {scheme}\nVorticity confinement: {vor_eps}")

    kw = dict(pressure="jacobi", n_iter=args.jacobi) if args.jacobi > 0 else {}
    if args.obstacle_image is not None:
        kw["obstacle_image"] = Path(args.obstacle_image)
    cls = FluidSimulator if args.no_dye else DyeFluidSimulator
    fluid_sim = cls.create(n_bc, resolution, dt, dx, re, vor_eps, scheme, **kw)

    if not args.no_graph:
        fluid_sim.enable_cuda_graph()
    out = Path(args.output)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for step in range(args.steps):
        fluid_sim.step()
        if args.dump_every and (step + 1) % args.dump_every == 0:
            out.mkdir(parents=True, exist_ok=True)
            np.savez(str(out / f"step_{step + 1:06}.npz"), **fluid_sim.field_to_numpy())
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    cells = 2 * resolution * resolution
    print(f"{args.steps} steps in {t:.3f} s: {args.steps / t:.1f} steps/s, {cells * args.steps / t / 1e9:.3f} G cell-updates/s")


if __name__ == "__main__":
    main()
