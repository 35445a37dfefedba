#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference source.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py [--only masks|kernels|traj]

The reference (`/root/reference/fs/*.py`) is imported as-is; its `import taichi` resolves to the
pure-Python semantics shim in `oracle/ti_shim/` (real Taichi cannot be installed here).
So the formulas, control flow, write masks, buffer swaps and scene builders recorded here are the
reference's own; only the fp32 lowering rules are ours (documented in the shim's docstring).

Outputs (committed):
    masks_small.npz      bc1..5 masks + bc_const at small resolutions (exact arrays)
    masks_sha256.json    sha256 of mask / bc_const at larger resolutions incl. the BASELINE configs
    kernels_r24.npz      one call of every hot-path reference kernel on seeded random fields
    traj_*.npz           N-step trajectories through the reference's own Solver.update()
"""
from __future__ import annotations

import argparse
import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parents[1]
sys.path.insert(0, str(REPO / "oracle" / "ti_shim"))
sys.path.insert(0, "/root/reference")

import taichi as ti  # noqa: E402  (the shim)
from fs.advection import advect_kk_scheme, advect_upwind  # noqa: E402
from fs.boundary_condition import (  # noqa: E402
    create_boundary_condition1,
    create_boundary_condition2,
    create_boundary_condition3,
    create_boundary_condition4,
    create_boundary_condition5,
    get_boundary_condition,
)
from fs.pressure_updater import JacobiPressureUpdater, RedBlackSorPressureUpdater  # noqa: E402
from fs.solver import CipMacSolver, DyeCipMacSolver, DyeMacSolver, MacSolver, clamp_field, limit_field  # noqa: E402
from fs.vorticity_confinement import VorticityConfinement  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bc_arrays(num: int, res: int):
    bc = get_boundary_condition(num, res, enable_dye=False)
    return bc, bc._bc_const.to_numpy(), bc._bc_mask.to_numpy()


# --------------------------------------------------------------------------- masks
def gen_masks() -> None:
    out = {}
    for num in (1, 2, 3, 4, 5):
        for res in (16, 24, 32, 40, 64):
            _, const, mask = bc_arrays(num, res)
            out[f"bc{num}_r{res}_mask"] = mask
            out[f"bc{num}_r{res}_const"] = const
    np.savez_compressed(HERE / "masks_small.npz", **out)

    shas = {}
    big = [(1, 128), (2, 128), (3, 128), (4, 128), (5, 128), (1, 256), (3, 256), (5, 400),
           (2, 2048), (5, 2048), (3, 1024), (2, 8192)]
    for num, res in big:
        t = time.time()
        _, const, mask = bc_arrays(num, res)
        shas[f"bc{num}_r{res}"] = {"mask": sha(mask), "const": sha(const),
                                   "fluid": int((mask == 0).sum())}
        print(f"bc{num} res{res} {time.time() - t:.1f}s", flush=True)
    (HERE / "masks_sha256.json").write_text(json.dumps(shas, indent=1))


# --------------------------------------------------------------------------- kernels
def gen_kernels(res: int = 24) -> None:
    rng = np.random.default_rng(20260925)
    out = {}
    dt = 0.05 / res
    dx = 1.0 / res
    re = 300.0
    for num in (1, 2, 3, 4, 5):
        bc, _, mask = bc_arrays(num, res)
        X, Y = mask.shape
        pre = f"bc{num}/"

        def rv(scale=1.0, ch=2):
            shp = (X, Y, ch) if ch else (X, Y)
            return (rng.uniform(-1, 1, shp) * scale).astype(np.float32)

        vcf = VorticityConfinement(bc, dt, dx, 5.0)
        jac = JacobiPressureUpdater(bc, dt, dx, 1)
        sor = RedBlackSorPressureUpdater(bc, dt, dx, 1.3, 1)
        cip = CipMacSolver(bc, jac, dt, dx, re, vcf)
        macs = {"upwind": MacSolver(bc, jac, advect_upwind, dt, dx, re, None),
                "kk": MacSolver(bc, jac, advect_kk_scheme, dt, dx, re, None)}

        def F(a):
            f = ti.Vector.field(a.shape[2], ti.f32, shape=a.shape[:2]) if a.ndim == 3 else ti.field(ti.f32, shape=a.shape)
            f.from_numpy(a)
            return f

        v, p = rv(), rv(ch=0)
        vx, vy = rv(0.1 / dx), rv(0.1 / dx)
        out[pre + "v"], out[pre + "p"], out[pre + "vx"], out[pre + "vy"] = v, p, vx, vy

        # velocity / pressure BC (in place)
        f = F(v); bc.set_velocity_boundary_condition(f); out[pre + "vel_bc"] = f.to_numpy()
        f = F(p); bc.set_pressure_boundary_condition(f); out[pre + "p_bc"] = f.to_numpy()

        # MAC fused update (upwind, kk): masked write into a pre-filled vn
        vn0 = rv()
        out[pre + "vn0"] = vn0
        for name, s in macs.items():
            f = F(vn0); s._update_velocities(f, F(v), F(p)); out[pre + f"mac_{name}"] = f.to_numpy()

        # CIP non-advection (+grad)
        fn = F(vn0); cip._non_advection_phase(fn, F(v), F(p)); out[pre + "nonadv"] = fn.to_numpy()
        vxn0, vyn0 = rv(), rv()
        out[pre + "vxn0"], out[pre + "vyn0"] = vxn0, vyn0
        fxn, fyn = F(vxn0), F(vyn0)
        cip._non_advection_phase_grad(fxn, fyn, F(vx), F(vy), F(v), fn)
        out[pre + "nonadv_gx"], out[pre + "nonadv_gy"] = fxn.to_numpy(), fyn.to_numpy()

        # CIP advection
        a, b, c = F(vn0), F(vxn0), F(vyn0)
        fv = F(v)
        cip._advection_phase(a, b, c, fv, F(vx), F(vy), fv)
        out[pre + "cip_f"], out[pre + "cip_fx"], out[pre + "cip_fy"] = a.to_numpy(), b.to_numpy(), c.to_numpy()

        # set_grad
        a, b = F(vxn0), F(vyn0)
        cip._set_grad(a, b, F(v))
        out[pre + "grad_x"], out[pre + "grad_y"] = a.to_numpy(), b.to_numpy()

        # vorticity confinement
        w0, wa0 = rv(ch=0), np.abs(rv(ch=0))
        out[pre + "w0"], out[pre + "wa0"] = w0, wa0
        vcf.vorticity.from_numpy(w0); vcf.vorticity_abs.from_numpy(wa0)
        vcf._calc_vorticity(F(v))
        out[pre + "vort"], out[pre + "vort_abs"] = vcf.vorticity.to_numpy(), vcf.vorticity_abs.to_numpy()
        f = F(vn0); vcf._add_vorticity(f, F(v)); out[pre + "vort_add"] = f.to_numpy()

        # pressure sweeps (no BC inside _update)
        pn0 = rv(ch=0)
        out[pre + "pn0"] = pn0
        f = F(pn0); jac._update(f, F(p), F(v)); out[pre + "jacobi"] = f.to_numpy()
        f = F(pn0); sor._update(f, F(p), F(v)); out[pre + "rbsor"] = f.to_numpy()

        # limiter
        f = F(v * 12.0); limit_field(f, 10.0); out[pre + "limit"] = f.to_numpy()
        print(f"kernels bc{num} done", flush=True)

    out["meta"] = np.array([res, dt, dx, re, 5.0], dtype=np.float64)
    np.savez_compressed(HERE / f"kernels_r{res}.npz", **out)


# --------------------------------------------------------------------------- trajectories
TRAJ = [
    # name, bc, res, dt, re, vc, scheme, pressure(kind, n_iter), steps, init
    ("A_cip_bc2_r32_jac4_vc5", 2, 32, None, 1e4, 5.0, "cip", ("jacobi", 4), 4, "zero"),
    ("B_cip_bc1_r32_jac3_novc_rand", 1, 32, None, 100.0, None, "cip", ("jacobi", 3), 3, "rand"),
    ("C_upwind_bc1_r32_jac5_vc5", 1, 32, 0.005, 100.0, 5.0, "upwind", ("jacobi", 5), 5, "zero"),
    ("D_kk_bc3_r40_sor2_vc2p5_rand", 3, 40, None, 1000.0, 2.5, "kk", ("rbsor", 2), 3, "rand"),
    ("E_cip_bc5_r32_sor2_vc5", 5, 32, None, 1e6, 5.0, "cip", ("rbsor", 2), 3, "zero"),
    ("F_cip_bc4_r24_jac5_vc10_rand", 4, 24, None, 1e4, 10.0, "cip", ("jacobi", 5), 3, "rand"),
    ("G_upwind_bc2_r24_jac2_novc_rand", 2, 24, None, 500.0, None, "upwind", ("jacobi", 2), 3, "rand"),
    ("H_cip_bc2_r16_jac1_vc5_rand", 2, 16, None, 1e4, 5.0, "cip", ("jacobi", 1), 4, "rand"),
]


def build_solver(num, res, dt, re, vc, scheme, pressure):
    dt = dt if dt is not None else 0.05 / res
    dx = 1.0 / res
    bc = get_boundary_condition(num, res, enable_dye=False)
    vcf = VorticityConfinement(bc, dt, dx, vc) if vc is not None else None
    if pressure[0] == "jacobi":
        pu = JacobiPressureUpdater(bc, dt, dx, pressure[1])
    else:
        pu = RedBlackSorPressureUpdater(bc, dt, dx, 1.3, pressure[1])
    if scheme == "cip":
        s = CipMacSolver(bc, pu, dt, dx, re, vcf)
    else:
        s = MacSolver(bc, pu, advect_upwind if scheme == "upwind" else advect_kk_scheme, dt, dx, re, vcf)
    return s, dt, dx


def state_of(s) -> dict:
    d = {"v_cur": s.v.current.to_numpy(), "v_nxt": s.v.next.to_numpy(),
         "p_cur": s.p.current.to_numpy(), "p_nxt": s.p.next.to_numpy()}
    if hasattr(s, "vx"):
        d.update(vx_cur=s.vx.current.to_numpy(), vx_nxt=s.vx.next.to_numpy(),
                 vy_cur=s.vy.current.to_numpy(), vy_nxt=s.vy.next.to_numpy())
    if s.vorticity_confinement is not None:
        d.update(vort=s.vorticity_confinement.vorticity.to_numpy(),
                 vort_abs=s.vorticity_confinement.vorticity_abs.to_numpy())
    return d


def gen_traj(only: str | None = None) -> None:
    for name, num, res, dt, re, vc, scheme, pressure, steps, init in TRAJ:
        if only and only not in name:
            continue
        t0 = time.time()
        s, dt_, dx_ = build_solver(num, res, dt, re, vc, scheme, pressure)
        X, Y = s.resolution
        out = {"meta_num": num, "meta_res": res, "meta_dt": dt_, "meta_dx": dx_, "meta_re": re,
               "meta_vc": -1.0 if vc is None else vc, "meta_scheme": scheme,
               "meta_pressure": pressure[0], "meta_n_iter": pressure[1], "meta_steps": steps}
        if init == "rand":
            rng = np.random.default_rng(sum(map(ord, name)))

            def rv(scale, ch):
                shp = (X, Y, ch) if ch else (X, Y)
                return (rng.uniform(-1, 1, shp) * scale).astype(np.float32)

            s.v.current.from_numpy(rv(0.5, 2)); s.v.next.from_numpy(rv(0.5, 2))
            s.p.current.from_numpy(rv(1.0, 0)); s.p.next.from_numpy(rv(1.0, 0))
            if hasattr(s, "vx"):
                for b in (s.vx, s.vy):
                    b.current.from_numpy(rv(0.05 / dx_, 2)); b.next.from_numpy(rv(0.05 / dx_, 2))
            if s.vorticity_confinement is not None:
                s.vorticity_confinement.vorticity.from_numpy(rv(1.0, 0))
                s.vorticity_confinement.vorticity_abs.from_numpy(np.abs(rv(1.0, 0)))
        for k, a in state_of(s).items():
            out[f"s0_{k}"] = a
        for n in range(1, steps + 1):
            s.update()
            for k, a in state_of(s).items():
                out[f"s{n}_{k}"] = a
        np.savez_compressed(HERE / f"traj_{name}.npz", **out)
        print(f"traj {name}: {time.time() - t0:.1f}s", flush=True)


# --------------------------------------------------------------------------- dye
DYE_TRAJ = [
    ("dyeA_cip_bc1_r24_sor2_vc5", 1, 24, None, 1e4, 5.0, "cip", ("rbsor", 2), 3, "zero"),
    ("dyeB_cip_bc2_r16_jac3_novc_rand", 2, 16, None, 300.0, None, "cip", ("jacobi", 3), 3, "rand"),
    ("dyeC_upwind_bc4_r24_jac2_vc5_rand", 4, 24, None, 500.0, 5.0, "upwind", ("jacobi", 2), 3, "rand"),
    ("dyeD_kk_bc5_r16_sor2_novc_rand", 5, 16, None, 1000.0, None, "kk", ("rbsor", 2), 3, "rand"),
    ("dyeE_cip_bc3_r20_jac4_vc5", 3, 20, None, 1e6, 5.0, "cip", ("jacobi", 4), 3, "zero"),
]


def dye_state_of(s) -> dict:
    d = state_of(s)
    d.update(dye_cur=s.dye.current.to_numpy(), dye_nxt=s.dye.next.to_numpy())
    if hasattr(s, "dyex"):
        d.update(dyex_cur=s.dyex.current.to_numpy(), dyex_nxt=s.dyex.next.to_numpy(),
                 dyey_cur=s.dyey.current.to_numpy(), dyey_nxt=s.dyey.next.to_numpy())
    return d


def gen_dye() -> None:
    # scene dye arrays
    out = {}
    for num in (1, 2, 3, 4, 5):
        for res in (16, 20, 24, 40):
            bc = get_boundary_condition(num, res, enable_dye=True)
            out[f"bc{num}_r{res}_dye"] = bc._bc_dye.to_numpy()
            out[f"bc{num}_r{res}_mask"] = bc._bc_mask.to_numpy()
            out[f"bc{num}_r{res}_const"] = bc._bc_const.to_numpy()
    np.savez_compressed(HERE / "dye_scenes.npz", **out)
    for name, num, res, dt, re, vc, scheme, pressure, steps, init in DYE_TRAJ:
        t0 = time.time()
        dt_ = dt if dt is not None else 0.05 / res
        dx_ = 1.0 / res
        bc = get_boundary_condition(num, res, enable_dye=True)
        vcf = VorticityConfinement(bc, dt_, dx_, vc) if vc is not None else None
        pu = (JacobiPressureUpdater(bc, dt_, dx_, pressure[1]) if pressure[0] == "jacobi"
              else RedBlackSorPressureUpdater(bc, dt_, dx_, 1.3, pressure[1]))
        if scheme == "cip":
            s = DyeCipMacSolver(bc, pu, dt_, dx_, re, vcf)
        else:
            s = DyeMacSolver(bc, pu, advect_upwind if scheme == "upwind" else advect_kk_scheme, dt_, dx_, re, vcf)
        X, Y = s.resolution
        out = {"meta_num": num, "meta_res": res, "meta_dt": dt_, "meta_dx": dx_, "meta_re": re,
               "meta_vc": -1.0 if vc is None else vc, "meta_scheme": scheme, "meta_pressure": pressure[0],
               "meta_n_iter": pressure[1], "meta_steps": steps}
        if init == "rand":
            rng = np.random.default_rng(sum(map(ord, name)))

            def rv(scale, ch):
                shp = (X, Y, ch) if ch else (X, Y)
                return (rng.uniform(-1, 1, shp) * scale).astype(np.float32)

            s.v.current.from_numpy(rv(0.5, 2)); s.v.next.from_numpy(rv(0.5, 2))
            s.p.current.from_numpy(rv(1.0, 0)); s.p.next.from_numpy(rv(1.0, 0))
            s.dye.current.from_numpy(np.abs(rv(1.0, 3))); s.dye.next.from_numpy(np.abs(rv(1.0, 3)))
            if hasattr(s, "vx"):
                for b in (s.vx, s.vy):
                    b.current.from_numpy(rv(0.05 / dx_, 2)); b.next.from_numpy(rv(0.05 / dx_, 2))
                for b in (s.dyex, s.dyey):
                    b.current.from_numpy(rv(0.05 / dx_, 3)); b.next.from_numpy(rv(0.05 / dx_, 3))
            if s.vorticity_confinement is not None:
                s.vorticity_confinement.vorticity.from_numpy(rv(1.0, 0))
                s.vorticity_confinement.vorticity_abs.from_numpy(np.abs(rv(1.0, 0)))
        for k, a in dye_state_of(s).items():
            out[f"s0_{k}"] = a
        for n in range(1, steps + 1):
            s.update()
            for k, a in dye_state_of(s).items():
                out[f"s{n}_{k}"] = a
        np.savez_compressed(HERE / f"traj_{name}.npz", **out)
        print(f"dye traj {name}: {time.time() - t0:.1f}s", flush=True)


# --------------------------------------------------------------------------- render kernels
def gen_render() -> None:
    from fs.fluid_simulator import DyeFluidSimulator

    rng = np.random.default_rng(77)
    out = {}
    for num, res in ((2, 16), (3, 20)):
        dt, dx = 0.05 / res, 1.0 / res
        bc = get_boundary_condition(num, res, enable_dye=True)
        pu = JacobiPressureUpdater(bc, dt, dx, 1)
        sim = DyeFluidSimulator(DyeCipMacSolver(bc, pu, dt, dx, 1e4, None))
        s = sim._solver
        X, Y = s.resolution
        v = (rng.uniform(-1, 1, (X, Y, 2)) * 3).astype(np.float32)
        p = (rng.uniform(-1, 1, (X, Y)) * 40).astype(np.float32)
        dye = rng.uniform(0, 1, (X, Y, 3)).astype(np.float32)
        s.v.current.from_numpy(v); s.p.current.from_numpy(p); s.dye.current.from_numpy(dye)
        pre = f"bc{num}_r{res}/"
        out[pre + "v"], out[pre + "p"], out[pre + "dye"] = v, p, dye
        out[pre + "norm"] = sim.get_norm_field().to_numpy()
        out[pre + "pressure"] = sim.get_pressure_field().to_numpy()
        out[pre + "vorticity"] = sim.get_vorticity_field().to_numpy()
        out[pre + "dye_img"] = sim.get_dye_field().to_numpy()
    np.savez_compressed(HERE / "render.npz", **out)
    print("render done", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--traj", default=None, help="substring filter for trajectory names")
    a = ap.parse_args()
    if a.only in (None, "masks"):
        gen_masks()
    if a.only in (None, "kernels"):
        gen_kernels()
    if a.only in (None, "traj"):
        gen_traj(a.traj)
    if a.only in (None, "dye"):
        gen_dye()
    if a.only in (None, "render"):
        gen_render()
