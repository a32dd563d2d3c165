"""One launch of every kernel that profiles/ documents, for `ncu` (OM4p25-shaped fields, 12 / 6 steps).

    ncu --set full --clock-control none --import-source on -o gpurun_out/<name> python tools/ncu_target.py [what ...]

``what`` picks the groups: steric (self-reference, supplied reference, global, thermo-, halosteric), variants (the
one-pass three-height kernel), elementwise (spice, density, reference state, delta_rho), strat (calc_n2).
A warm-up call of each runs first so that the profiled launch is not the one that pays for module loading.
"""
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from momlevel_b200 import core, synth  # noqa: E402

what = set(sys.argv[1:]) or {"steric", "variants", "elementwise", "strat"}
nt, nz, ny, nx = 12, 75, 1080, 1440
grid = synth.make_grid(nz, ny, nx, seed=123, device="cuda")
T, S, V = synth.make_fields(grid, nt, seed=123, dtype=torch.float32)
pres = (grid["z_l"] * 1e4 + 101325.0).contiguous()
z_i, depth, z_l = grid["z_i"].contiguous(), grid["deptho"].contiguous(), grid["z_l"].contiguous()
rho_ref, _ = core.reference_state(T[0], S[0], V, pres)
half = nt // 2
calls = []
if "steric" in what:
    calls += [lambda: core.steric_local_selfref(T, S, V, z_i, depth, pres, want_rho_ref=False),
              lambda: core.steric_local(T, S, rho_ref, V, z_i, depth, pres),
              lambda: core.steric_global(T, S, V, pres),
              lambda: core.steric_local(T, S[0], rho_ref, V, z_i, depth, pres, s_bcast=True),
              lambda: core.steric_local(T[0], S, rho_ref, V, z_i, depth, pres, t_bcast=True)]
if "variants" in what:
    calls += [lambda: core.steric_local_variants(T, S, V, z_i, depth, pres, want_rho_ref=False)]
if "elementwise" in what:
    calls += [lambda: core.flament_spice(T[:half], S[:half]),
              lambda: core.eos_eval("Wright", "density", T[:half], S[:half], pres, z_axis=1),
              lambda: core.reference_state(T[0], S[0], V, pres),
              lambda: core.delta_rho(T[:half], S[:half], rho_ref, V, pres)]
if "strat" in what:
    calls += [lambda: core.calc_n2(T[:half], S[:half], z_l)]
for rep in range(2):  # ncu: --launch-skip the first half
    for fn in calls:
        r = fn()
        torch.cuda.synchronize()
        del r
print("launches", core.launch_count())
